/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's oibvh collision path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker / reported baseline. The product (oibvh_b200/csrc, liboibvh_b200.so)
 * never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED. Every function here is checked against the UNMODIFIED reference CPU code
 * (oracle/_ref/liboibvh_ref.so, built by oracle/Makefile from /root/reference) in tests/test_oracle_vs_ref.py,
 * and against golden vectors frozen from that code in tests/golden/ (generator: tools/make_golden.py).
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 * Compile with -ffp-contract=off: the reference CPU path is plain IEEE fp32 without fused multiply-add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint32_t u32;
typedef uint64_t u64;

/* ------------------------------------------------------------------------------------------------
 * integer helpers -- include/cuda/utils.cuh:66-86
 * ---------------------------------------------------------------------------------------------- */
u32 orc_next_pow2(u32 x)
{
    x--;
    x |= x >> 1;
    x |= x >> 2;
    x |= x >> 4;
    x |= x >> 8;
    x |= x >> 16;
    x++;
    return x;
}

u32 orc_ilog2(u32 x) /* floor(log2 x); x = 0 gives 2^32-1 like the reference's 31 - clz(0)=32 wrap */
{
    if (x == 0) return (u32)-1;
    return 31u - (u32)__builtin_clz(x);
}

static u32 popc(u32 x) { return (u32)__builtin_popcount(x); }

/* shifts by >= 32 are UB in C; the reference relies on the device/x86 behaviour only for the root
 * (level-1 = 2^32-1) which every caller below special-cases. */
static u32 shr(u32 v, u32 s) { return s >= 32 ? 0u : (v >> s); }

/* ------------------------------------------------------------------------------------------------
 * oibvh layout -- include/cuda/oibvh.cuh:56-182
 * ---------------------------------------------------------------------------------------------- */
u32 orc_get_size(u32 t) { return 2 * t + popc(orc_next_pow2(t) - t) - 1; } /* oibvh.cuh:56-59 */

u32 orc_level_virtual_count(u32 li, u32 lli, u32 vl) { return shr(vl, lli - li); } /* :68-72 */

u32 orc_level_real_count(u32 li, u32 lli, u32 vl) /* :81-85 */
{
    return (1u << li) - orc_level_virtual_count(li, lli, vl);
}

u32 orc_level_all_virtual_count(u32 li, u32 lli, u32 vl) /* :94-99 */
{
    const u32 v = orc_level_virtual_count(li, lli, vl);
    return (v << 1) - popc(v);
}

u32 orc_implicit_to_real(u32 implicitIdx, u32 leafLev, u32 vl) /* :108-113 */
{
    const u32 level = orc_ilog2(implicitIdx + 1);
    if (level == 0) return implicitIdx; /* root: no virtual nodes above it */
    return implicitIdx - orc_level_all_virtual_count(level - 1, leafLev, vl);
}

u32 orc_real_to_implicit(u32 realIdx, u32 leafLev, u32 vl) /* :122-136 */
{
    if (realIdx == 0) return 0;
    const u32 level = orc_ilog2(orc_next_pow2(realIdx)) - 1;
    const u32 levelAllVirtual = orc_level_all_virtual_count(level, leafLev, vl);
    const u32 levelAllReal = (1u << (level + 1)) - 2 - levelAllVirtual;
    if (levelAllReal < realIdx) return realIdx + levelAllVirtual;
    return realIdx + levelAllVirtual - orc_level_virtual_count(level, leafLev, vl);
}

int orc_have_rchild(u32 implicitIdx, u32 leafLev, u32 vl) /* :145-157 */
{
    const u32 nextLevel = orc_ilog2(implicitIdx + 1) + 1;
    return 2 * implicitIdx + 4 <= (1u << nextLevel) + orc_level_real_count(nextLevel, leafLev, vl);
}

u32 orc_most_left_descendant(u32 implicitIdx, u32 descendLev) /* :165-169 */
{
    return (1u << descendLev) * implicitIdx + (1u << descendLev) - 1;
}

u32 orc_most_right_valid(u32 level, u32 leafLev, u32 vl) /* :178-182 */
{
    return ((1u << (level + 1)) - 2) - orc_level_virtual_count(level, leafLev, vl);
}

/* ------------------------------------------------------------------------------------------------
 * glm / thrust scalar semantics
 *   glm::min(x,y) = (y < x) ? y : x ; glm::max(x,y) = (x < y) ? y : x   third/glm/detail/func_common.inl:17-30
 *   thrust::min(l,r) = r < l ? r : l ; thrust::max(l,r) = l < r ? r : l  (CUDA thrust/detail/minmax.h)
 * ---------------------------------------------------------------------------------------------- */
static inline float gmin(float x, float y) { return (y < x) ? y : x; }
static inline float gmax(float x, float y) { return (x < y) ? y : x; }

/* Mesh::setupAABB -- src/utils/mesh.cpp:91-98. Starts from (+FLT_MAX, -FLT_MAX) (utils.h:17-18);
 * note the argument order: glm::max(vertex, current), glm::min(vertex, current). */
void orc_mesh_aabb(const float* pos, u32 V, float* out6)
{
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (u32 i = 0; i < V; i++)
        for (int a = 0; a < 3; a++)
        {
            const float p = pos[3 * (size_t)i + a];
            mx[a] = gmax(p, mx[a]);
            mn[a] = gmin(p, mn[a]);
        }
    memcpy(out6, mn, 12);
    memcpy(out6 + 3, mx, 12);
}

/* leaf box of one face -- src/cuda/oibvh.cu:33-39 (glm::min(glm::min(v0,v1),v2)), identical values to
 * SimpleBVH's merge chain src/cpu/simpleBVH.cpp:136-139 / 100-103 */
static inline void face_aabb(const float* pos, const u32* f, float* out6)
{
    for (int a = 0; a < 3; a++)
    {
        const float v0 = pos[3 * (size_t)f[0] + a], v1 = pos[3 * (size_t)f[1] + a], v2 = pos[3 * (size_t)f[2] + a];
        out6[a] = gmin(gmin(v0, v1), v2);
        out6[3 + a] = gmax(gmax(v0, v1), v2);
    }
}

void orc_leaf_aabbs(const float* pos, const u32* faces, u32 T, float* out)
{
    for (u32 i = 0; i < T; i++) face_aabb(pos, faces + 3 * (size_t)i, out + 6 * (size_t)i);
}

/* ------------------------------------------------------------------------------------------------
 * Morton keys -- src/cuda/oibvh.cu:6-22 (strech_by_3, morton3D), :63-69 (normalisation)
 * ---------------------------------------------------------------------------------------------- */
static inline u32 spread3(u32 x) /* oibvh.cu:6-15 */
{
    x = x & 0x3ffu;
    x = (x | (x << 16)) & 0x30000ffu;
    x = (x | (x << 8)) & 0x300f00fu;
    x = (x | (x << 4)) & 0x30c30c3u;
    x = (x | (x << 2)) & 0x9249249u;
    return x;
}

static inline u32 quantise(float q) /* (unsigned)thrust::min(thrust::max(q*1024, 0), 1023), oibvh.cu:19-21 */
{
    float s = q * 1024.0f;
    float a = (s < 0.0f) ? 0.0f : s;       /* thrust::max(s, 0): lhs<rhs ? rhs : lhs ; NaN stays NaN */
    float b = (1023.0f < a) ? 1023.0f : a; /* thrust::min(a, 1023): rhs<lhs ? rhs : lhs ; NaN stays NaN */
    if (b != b) return 0u;                 /* (unsigned)NaN: CUDA cvt.rzi.u32.f32 gives 0 (SURVEY.md §2.2 hazards) */
    return (u32)b;
}

u32 orc_morton_key(const float* box6, const float* mesh_aabb6)
{
    u32 u[3];
    for (int a = 0; a < 3; a++)
    {
        const float centroid = (box6[a] + box6[3 + a]) * 0.5f; /* oibvh.cu:63 */
        const float offset = centroid - mesh_aabb6[a];         /* :64 */
        const float length = mesh_aabb6[3 + a] - mesh_aabb6[a]; /* :65 */
        u[a] = quantise(offset / length);                      /* :68 */
    }
    return spread3(u[0]) << 2 | spread3(u[1]) << 1 | spread3(u[2]);
}

void orc_morton_keys(const float* pos, const u32* faces, u32 T, const float* mesh_aabb6, u32* keys)
{
    for (u32 i = 0; i < T; i++)
    {
        float box[6];
        face_aabb(pos, faces + 3 * (size_t)i, box);
        keys[i] = orc_morton_key(box, mesh_aabb6);
    }
}

/* thrust::stable_sort_by_key (src/cuda/oibvhTree.cu:295-296): stable ascending by key; ties keep input order.
 * Any stable sort is equivalent; this is a 4x8-bit LSD counting sort producing perm[sorted] = original. */
void orc_stable_sort_perm(const u32* keys, u32 T, u32* perm)
{
    u32* ka = (u32*)malloc(sizeof(u32) * (size_t)T);
    u32* kb = (u32*)malloc(sizeof(u32) * (size_t)T);
    u32* pb = (u32*)malloc(sizeof(u32) * (size_t)T);
    u32* pa = perm;
    memcpy(ka, keys, sizeof(u32) * (size_t)T);
    for (u32 i = 0; i < T; i++) pa[i] = i;
    for (int pass = 0; pass < 4; pass++)
    {
        size_t hist[257] = {0};
        const int sh = 8 * pass;
        for (u32 i = 0; i < T; i++) hist[((ka[i] >> sh) & 255u) + 1]++;
        for (int d = 0; d < 256; d++) hist[d + 1] += hist[d];
        for (u32 i = 0; i < T; i++)
        {
            const size_t dst = hist[(ka[i] >> sh) & 255u]++;
            kb[dst] = ka[i];
            pb[dst] = pa[i];
        }
        u32* t = ka; ka = kb; kb = t;
        t = pa; pa = pb; pb = t;
    }
    /* an even number of swaps: ka/pa are the original buffers again, i.e. pa == perm holds the result */
    free(ka);
    free(kb);
    free(pb);
}

/* ------------------------------------------------------------------------------------------------
 * Tree construction -- leaves (oibvh.cu:24-40 / :42-62 written at aabbs + internalCount, oibvhTree.cu:211-217,
 * 260-266) then bottom-up parent = left U right, or = left when the right child is virtual
 * (oibvh.cu:72-78 merge_aabb = glm::min/max(left,right); :179-187, :204-211 choice), every node stored at
 * implicit_to_real(i). The reference kernels climb in 256-wide groups (oibvhTree.cu:126-155); the result is
 * schedule-independent, so this walks whole levels.
 * nodes: N x 6 floats (min.xyz, max.xyz) in real-index order, N = orc_get_size(T). faces are taken in the
 * order given (Morton-sorted for the GPU tree, input order for the SimpleBVH comparison).
 * ---------------------------------------------------------------------------------------------- */
void orc_tree_from_faces(const float* pos, const u32* faces, u32 T, float* nodes)
{
    const u32 P = orc_next_pow2(T);
    const u32 leafLev = orc_ilog2(P);
    const u32 vl = P - T;
    const u32 N = orc_get_size(T);
    float* leaves = nodes + 6 * (size_t)(N - T);
    for (u32 i = 0; i < T; i++) face_aabb(pos, faces + 3 * (size_t)i, leaves + 6 * (size_t)i);
    for (u32 lev = leafLev; lev-- > 0;)
    {
        const u32 first = (1u << lev) - 1;
        const u32 last = orc_most_right_valid(lev, leafLev, vl);
        for (u32 i = first; i <= last; i++)
        {
            const u32 r = orc_implicit_to_real(i, leafLev, vl);
            const u32 lc = orc_implicit_to_real(2 * i + 1, leafLev, vl);
            float* dst = nodes + 6 * (size_t)r;
            const float* L = nodes + 6 * (size_t)lc;
            if (orc_have_rchild(i, leafLev, vl))
            {
                const float* R = L + 6; /* right child real index = left + 1, oibvh.cu:182 */
                for (int a = 0; a < 3; a++)
                {
                    dst[a] = gmin(L[a], R[a]);
                    dst[3 + a] = gmax(L[3 + a], R[3 + a]);
                }
            }
            else
                memcpy(dst, L, 24);
        }
    }
}

/* OibvhTree::build -- src/cuda/oibvhTree.cu:237-388: keys, stable sort carrying the faces, tree over the sorted
 * faces. Outputs: nodes (N x 6), sorted_faces (T x 3), perm (T, sorted position -> original face id; the
 * reference does not keep it, SURVEY.md §3.1, it is needed to canonicalise pair sets). */
void orc_build(const float* pos, const u32* faces, u32 T, const float* mesh_aabb6, float* nodes, u32* sorted_faces,
               u32* perm, u32* sorted_keys /* optional */)
{
    u32* keys = (u32*)malloc(sizeof(u32) * (size_t)T);
    orc_morton_keys(pos, faces, T, mesh_aabb6, keys);
    orc_stable_sort_perm(keys, T, perm);
    for (u32 i = 0; i < T; i++)
    {
        memcpy(sorted_faces + 3 * (size_t)i, faces + 3 * (size_t)perm[i], 12);
        if (sorted_keys) sorted_keys[i] = keys[perm[i]];
    }
    free(keys);
    orc_tree_from_faces(pos, sorted_faces, T, nodes);
}

/* OibvhTree::refit -- src/cuda/oibvhTree.cu:193-235: same tree over the (already sorted) faces and new positions */
void orc_refit(const float* pos, const u32* sorted_faces, u32 T, float* nodes)
{
    orc_tree_from_faces(pos, sorted_faces, T, nodes);
}

/* ------------------------------------------------------------------------------------------------
 * Narrow phase -- src/utils/utils.cpp:71-169 (== third/gProximity/cuda_intersect_tritri.h:248-267, 350-434
 * when not FMA-contracted): 17-axis SAT after translating P1 to the origin.
 *   glm::cross(x,y) = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)  func_geometric.inl:74-77
 *   glm::dot(a,b)   = (a.x*b.x + a.y*b.y) + a.z*b.z                              func_geometric.inl:52-53
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
    float x, y, z;
} v3;

static inline v3 vsub(v3 a, v3 b)
{
    v3 r = {a.x - b.x, a.y - b.y, a.z - b.z};
    return r;
}
static inline v3 vcross(v3 x, v3 y)
{
    v3 r = {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
    return r;
}
static inline float vdot(v3 a, v3 b)
{
    const float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return tx + ty + tz;
}

static inline int project6(v3 ax, v3 p1, v3 p2, v3 p3, v3 q1, v3 q2, v3 q3) /* utils.cpp:71-95 */
{
    const float P1 = vdot(ax, p1), P2 = vdot(ax, p2), P3 = vdot(ax, p3);
    const float Q1 = vdot(ax, q1), Q2 = vdot(ax, q2), Q3 = vdot(ax, q3);
    const float mx1 = fmaxf(fmaxf(P1, P2), P3);
    const float mn1 = fminf(fminf(P1, P2), P3);
    const float mx2 = fmaxf(fmaxf(Q1, Q2), Q3);
    const float mn2 = fminf(fminf(Q1, Q2), Q3);
    if (mn1 > mx2) return 0;
    if (mn2 > mx1) return 0;
    return 1;
}

int orc_tri_tri(const float* p, const float* q) /* utils.cpp:97-169 */
{
    const v3 P1 = {p[0], p[1], p[2]}, P2 = {p[3], p[4], p[5]}, P3 = {p[6], p[7], p[8]};
    const v3 Q1 = {q[0], q[1], q[2]}, Q2 = {q[3], q[4], q[5]}, Q3 = {q[6], q[7], q[8]};
    const v3 p1 = {0.0f, 0.0f, 0.0f};
    const v3 p2 = vsub(P2, P1), p3 = vsub(P3, P1);
    const v3 q1 = vsub(Q1, P1), q2 = vsub(Q2, P1), q3 = vsub(Q3, P1);
    const v3 e1 = vsub(p2, p1), e2 = vsub(p3, p2), e3 = vsub(p1, p3);
    const v3 f1 = vsub(q2, q1), f2 = vsub(q3, q2), f3 = vsub(q1, q3);
    const v3 n1 = vcross(e1, e2), m1 = vcross(f1, f2);
    const v3 e[3] = {e1, e2, e3}, f[3] = {f1, f2, f3};
    if (!project6(n1, p1, p2, p3, q1, q2, q3)) return 0;
    if (!project6(m1, p1, p2, p3, q1, q2, q3)) return 0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            if (!project6(vcross(e[i], f[j]), p1, p2, p3, q1, q2, q3)) return 0; /* ef11..ef33 */
    for (int i = 0; i < 3; i++)
        if (!project6(vcross(e[i], n1), p1, p2, p3, q1, q2, q3)) return 0; /* g1..g3 */
    for (int j = 0; j < 3; j++)
        if (!project6(vcross(f[j], m1), p1, p2, p3, q1, q2, q3)) return 0; /* h1..h3 */
    return 1;
}

/* aabb_box_t::overlap -- include/utils/utils.h:39-44 (inclusive) */
int orc_aabb_overlap(const float* a, const float* b)
{
    return (a[0] <= b[3] && a[3] >= b[0]) && (a[1] <= b[4] && a[4] >= b[1]) && (a[2] <= b[5] && a[5] >= b[2]);
}

/* ------------------------------------------------------------------------------------------------
 * Collision detection -- SimpleCollide::detect, src/cpu/simpleCollide.cpp:47-160, restated over array trees:
 * DFS with an explicit stack from the root pair of every object pair i<j; overlap -> both leaves: SAT, record
 * {i, j, triA, triB}; otherwise each side contributes {left[, right]} (or itself when it is a leaf).
 * Triangle ids are positions in the face order the tree was built over (sorted positions for a Morton tree).
 * level_hist (optional, 64 entries) counts BVTT nodes popped per depth like `a[depth]` (:63-68).
 * Returns the number of intersecting pairs; writes at most `capacity` records of 4 x u32.
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
    u32 a, b, depth;
} bvtt_t;

u64 orc_detect(u32 n_obj, const float* const* nodes, const u32* const* faces, const float* const* pos, const u32* T,
               u32* out_pairs, u64 capacity, u64* n_candidates, u64* level_hist)
{
    u64 n_pairs = 0, n_cand = 0;
    size_t cap = 1024, sp = 0;
    bvtt_t* stack = (bvtt_t*)malloc(sizeof(bvtt_t) * cap);
    if (level_hist) memset(level_hist, 0, 64 * sizeof(u64));
    for (u32 oi = 0; oi < n_obj; oi++)
        for (u32 oj = oi + 1; oj < n_obj; oj++)
        {
            const u32 PA = orc_next_pow2(T[oi]), PB = orc_next_pow2(T[oj]);
            const u32 LA = orc_ilog2(PA), LB = orc_ilog2(PB);
            const u32 vlA = PA - T[oi], vlB = PB - T[oj];
            const u32 leafBaseA = (1u << LA) - 1, leafBaseB = (1u << LB) - 1;
            sp = 0;
            stack[sp++] = (bvtt_t){0, 0, 0};
            while (sp)
            {
                const bvtt_t n = stack[--sp];
                if (level_hist && n.depth < 64) level_hist[n.depth]++;
                const float* A = nodes[oi] + 6 * (size_t)orc_implicit_to_real(n.a, LA, vlA);
                const float* B = nodes[oj] + 6 * (size_t)orc_implicit_to_real(n.b, LB, vlB);
                if (!orc_aabb_overlap(A, B)) continue;
                const int leafA = n.a >= leafBaseA, leafB = n.b >= leafBaseB;
                if (leafA && leafB)
                {
                    n_cand++;
                    const u32 ta = n.a - leafBaseA, tb = n.b - leafBaseB;
                    const u32* fa = faces[oi] + 3 * (size_t)ta;
                    const u32* fb = faces[oj] + 3 * (size_t)tb;
                    float p[9], q[9];
                    for (int k = 0; k < 3; k++)
                    {
                        memcpy(p + 3 * k, pos[oi] + 3 * (size_t)fa[k], 12);
                        memcpy(q + 3 * k, pos[oj] + 3 * (size_t)fb[k], 12);
                    }
                    if (orc_tri_tri(p, q))
                    {
                        if (n_pairs < capacity)
                        {
                            u32* o = out_pairs + 4 * n_pairs;
                            o[0] = oi; o[1] = oj; o[2] = ta; o[3] = tb;
                        }
                        n_pairs++;
                    }
                    continue;
                }
                u32 ca[2], cb[2], na = 0, nb = 0;
                if (leafA) ca[na++] = n.a;
                else
                {
                    ca[na++] = 2 * n.a + 1;
                    if (orc_have_rchild(n.a, LA, vlA)) ca[na++] = 2 * n.a + 2;
                }
                if (leafB) cb[nb++] = n.b;
                else
                {
                    cb[nb++] = 2 * n.b + 1;
                    if (orc_have_rchild(n.b, LB, vlB)) cb[nb++] = 2 * n.b + 2;
                }
                if (sp + 4 > cap)
                {
                    cap *= 2;
                    stack = (bvtt_t*)realloc(stack, sizeof(bvtt_t) * cap);
                }
                for (u32 i = 0; i < na; i++)
                    for (u32 j = 0; j < nb; j++) stack[sp++] = (bvtt_t){ca[i], cb[j], n.depth + 1};
            }
        }
    free(stack);
    if (n_candidates) *n_candidates = n_cand;
    return n_pairs;
}

/* Brute-force candidate set definition (SURVEY.md Appendix A "Broad phase"): {(a,b): AABB(tri a) overlaps
 * AABB(tri b)} -- used by small tests to check the traversal itself. Returns count, writes <= capacity pairs. */
u64 orc_candidates_bruteforce(const float* leavesA, u32 TA, const float* leavesB, u32 TB, u32* out2, u64 capacity)
{
    u64 n = 0;
    for (u32 i = 0; i < TA; i++)
        for (u32 j = 0; j < TB; j++)
            if (orc_aabb_overlap(leavesA + 6 * (size_t)i, leavesB + 6 * (size_t)j))
            {
                if (n < capacity)
                {
                    out2[2 * n] = i;
                    out2[2 * n + 1] = j;
                }
                n++;
            }
    return n;
}

/* UV sphere of the survey's known-answer table: vertices j = 0..n, i = 0..n-1 at
 * theta = float(pi)*j/n, phi = 2*float(pi)*i/n, p = (sin t cos p, cos t, sin t sin p) in fp32 with libm sinf/cosf;
 * faces (a,b,c),(b,d,c). pos: (n+1)*n*3 floats, faces: 2*n*n*3 u32. Not reference code -- a test input. */
void orc_gen_uv_sphere(u32 n, float* pos, u32* faces)
{
    const float pi = 3.14159265358979323846f;
    for (u32 j = 0; j <= n; j++)
        for (u32 i = 0; i < n; i++)
        {
            const float th = pi * (float)j / (float)n, ph = 2.0f * pi * (float)i / (float)n;
            float* p = pos + 3 * (size_t)(j * n + i);
            p[0] = sinf(th) * cosf(ph);
            p[1] = cosf(th);
            p[2] = sinf(th) * sinf(ph);
        }
    for (u32 j = 0; j < n; j++)
        for (u32 i = 0; i < n; i++)
        {
            const u32 a = j * n + i, b = j * n + (i + 1) % n, c = (j + 1) * n + i, d = (j + 1) * n + (i + 1) % n;
            u32* f = faces + 6 * (size_t)(j * n + i);
            f[0] = a; f[1] = b; f[2] = c;
            f[3] = b; f[4] = d; f[5] = c;
        }
}

/* transform_vec4_kernel -- src/cuda/transform.cu:35-40 with glm mat4*vec4 (third/glm/detail/type_mat4x4.inl:561-572):
 * (m0*x + m1*y) + (m2*z + m3*w), w = 1 (mesh.cpp:193), column-major M. In place on packed xyz. */
void orc_transform_positions(float* pos, u32 V, const float* M)
{
    for (u32 i = 0; i < V; i++)
    {
        float* p = pos + 3 * (size_t)i;
        const float x = p[0], y = p[1], z = p[2];
        for (int r = 0; r < 3; r++)
        {
            const float mul0 = M[0 + r] * x, mul1 = M[4 + r] * y, add0 = mul0 + mul1;
            const float mul2 = M[8 + r] * z, mul3 = M[12 + r] * 1.0f, add1 = mul2 + mul3;
            p[r] = add0 + add1;
        }
    }
}
