"""Thread-level transcription (pure Python, slow, small inputs) of the EXPERIMENTAL kernels in
oibvh_b200/csrc/sort_msd.cu: msd_plan_kernel and msd_sort_kernel, index formula by index formula (threads, warps,
items, slots, counts matrix, row scan, range-local passes). It exists because those kernels were written after the
round's GPU budget was spent: the transcription is checked against tools/sort_model.py and numpy's stable argsort in
tests/test_sort_model.py, so that the first GPU run starts from logic that is known to be right."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sort_model  # noqa: E402,F401

CAP, WIN, BINS, PLAN_THREADS = 8192, 4096, 65536, 1024
PER = BINS // PLAN_THREADS
THREADS, WARPS, IPT_MAX = 512, 16, 16

def plan_kernel(hist, T):
    """thread-level transcription of msd_plan_kernel"""
    sums = np.array([hist[t*PER:(t+1)*PER].sum() for t in range(PLAN_THREADS)])
    seg_start = np.cumsum(sums) - sums
    total = sums.sum(); max_bin = hist.max()
    last_start = np.zeros(PLAN_THREADS, np.int64); last_count = np.zeros(PLAN_THREADS, np.int64)
    for t in range(PLAN_THREADS):
        run = seg_start[t]; ls = lc = 0
        for k in range(PER):
            c = hist[t*PER+k]
            if c: ls, lc = run, c
            run += c
        last_start[t], last_count[t] = ls, lc
    carry = []
    for t in range(PLAN_THREADS):
        hp = False; ph = False; pw = 0
        for u in range(t-1, -1, -1):
            if last_count[u]:
                hp = True; ph = last_count[u] > CAP - WIN; pw = last_start[u] // WIN; break
        carry.append((hp, ph, pw))
    opens = np.zeros(PLAN_THREADS, np.int64)
    for t in range(PLAN_THREADS):
        run = seg_start[t]; hp, ph, pw = carry[t]
        for k in range(PER):
            c = hist[t*PER+k]
            if c:
                heavy = c > CAP - WIN; win = run // WIN
                if (not hp) or heavy or ph or win != pw: opens[t] += 1
                hp, ph, pw = True, heavy, win
            run += c
    first_range = np.cumsum(opens) - opens; n_ranges = opens.sum()
    range_of_bin = np.zeros(BINS, np.int64); range_start = np.zeros(1025, np.int64)
    for t in range(PLAN_THREADS):
        run = seg_start[t]; nxt = first_range[t]; hp, ph, pw = carry[t]
        for k in range(PER):
            c = hist[t*PER+k]
            idv = nxt - 1 if nxt else 0
            if c:
                heavy = c > CAP - WIN; win = run // WIN
                if (not hp) or heavy or ph or win != pw:
                    if nxt < 1024: range_start[nxt] = run
                    nxt += 1
                idv = nxt - 1
                hp, ph, pw = True, heavy, win
            range_of_bin[t*PER+k] = min(idv, 1023)
            run += c
    fallback = max_bin > CAP or n_ranges > 1024 or total != T or n_ranges == 0
    if n_ranges <= 1024: range_start[n_ranges] = T
    return n_ranges, fallback, range_of_bin, range_start

def msd_sort_kernel(keys_a, G, n_ranges, range_of_bin, range_start):
    T = len(keys_a)
    keys_b = np.zeros(T, np.uint32); vals_b = np.zeros(T, np.uint32)
    out_k = np.zeros(T, np.uint32); out_v = np.zeros(T, np.uint32)
    ipt = (T + G*THREADS - 1)//(G*THREADS); assert ipt <= IPT_MAX
    chunk = THREADS*ipt
    mat = np.zeros((1024, G), np.int64)
    cta_state = []
    # ---- P1 up to grid barrier 1
    for cta in range(G):
        cta_base = cta*chunk; cta_valid = min(chunk, T-cta_base) if cta_base < T else 0
        cnt = np.zeros((WARPS, 1024), np.int64)
        key = np.full((WARPS, IPT_MAX, 32), 0xffffffff, np.uint64); rid = np.zeros((WARPS, IPT_MAX, 32), np.int64)
        rank = np.zeros((WARPS, IPT_MAX, 32), np.int64); valid = np.zeros((WARPS, IPT_MAX, 32), bool)
        for w in range(WARPS):
            warp_base = cta_base + w*(32*ipt)
            for j in range(ipt):
                for l in range(32):
                    i = warp_base + j*32 + l
                    if i < T:
                        valid[w,j,l] = True; key[w,j,l] = keys_a[i]; rid[w,j,l] = range_of_bin[int(keys_a[i]) >> 14]
            for j in range(ipt):
                for l in range(32):
                    if not valid[w,j,l]: continue
                    r = rid[w,j,l]
                    peers = [m for m in range(32) if valid[w,j,m] and rid[w,j,m] == r]
                    lower = sum(1 for m in peers if m < l)
                    rank[w,j,l] = cnt[w, r] + lower          # prev (leader's read) + lower
                for r in set(rid[w,j][valid[w,j]].tolist()):
                    cnt[w, r] += int(((rid[w,j] == r) & valid[w,j]).sum())
        tot = np.zeros(1024, np.int64)
        for r in range(1024):
            run = 0
            for w in range(WARPS):
                c = cnt[w, r]; cnt[w, r] = run; run += c
            tot[r] = run
            if r < n_ranges: mat[r, cta] = run
        base = np.cumsum(tot) - tot
        kv = [None]*max(cta_valid, 1)
        for w in range(WARPS):
            warp_base = cta_base + w*(32*ipt)
            for j in range(ipt):
                for l in range(32):
                    if valid[w,j,l]:
                        r = rid[w,j,l]; i = warp_base + j*32 + l
                        slot = base[r] + cnt[w, r] + rank[w,j,l]
                        assert kv[slot] is None
                        kv[slot] = (int(key[w,j,l]), i | (int(r) << 21))
        cta_state.append((cta_valid, base, kv))
    # ---- row scan
    matp = np.cumsum(mat, axis=1) - mat
    # ---- write-out
    for cta in range(G):
        cta_valid, base, kv = cta_state[cta]
        gb = np.array([range_start[r] + matp[r, cta] - base[r] if r < n_ranges else 0 for r in range(1024)])
        for s in range(cta_valid):
            k, y = kv[s]; dst = gb[y >> 21] + s
            keys_b[dst] = k; vals_b[dst] = y & ((1 << 21) - 1)
    # ---- P2
    for r in range(n_ranges):
        a = int(range_start[r]); n = min(int(range_start[r+1]) - a, CAP)
        lkey = keys_b[a:a+n].copy(); ids = np.arange(n)
        wchunk = (((n + WARPS - 1)//WARPS) + 31) & ~31
        for p in range(4):
            shift = 8*p
            tab = np.zeros((WARPS, 256), np.int64); rk = {}
            for w in range(WARPS):
                for s in range(IPT_MAX):
                    if s*32 >= wchunk: break
                    lanes = [(l, w*wchunk + s*32 + l) for l in range(32) if w*wchunk + s*32 + l < n]
                    dg = {l: (int(lkey[ids[i]]) >> shift) & 255 for l, i in lanes}
                    for l, i in lanes:
                        d = dg[l]; lower = sum(1 for m, _ in lanes if m < l and dg[m] == d)
                        rk[(w, s, l)] = tab[w, d] + lower
                    for d in set(dg.values()):
                        tab[w, d] += sum(1 for v in dg.values() if v == d)
            tot = np.zeros(256, np.int64)
            for d in range(256):
                run = 0
                for w in range(WARPS):
                    c = tab[w, d]; tab[w, d] = run; run += c
                tot[d] = run
            dbase = np.cumsum(tot) - tot
            dst = np.full(n, -1)
            for w in range(WARPS):
                for s in range(IPT_MAX):
                    if s*32 >= wchunk: break
                    for l in range(32):
                        i = w*wchunk + s*32 + l
                        if i < n:
                            idx = ids[i]; d = (int(lkey[idx]) >> shift) & 255
                            slot = dbase[d] + tab[w, d] + rk[(w, s, l)]
                            assert dst[slot] == -1
                            dst[slot] = idx
            ids = dst
        out_k[a:a+n] = lkey[ids]; out_v[a:a+n] = vals_b[a + ids]
    return out_k, out_v


def sort(keys, G=5):
    """-> (sorted keys, permutation) through the transcribed kernels, or None when the plan asks for the fallback"""
    keys = np.asarray(keys, np.uint32)
    hist = np.bincount(keys >> 14, minlength=BINS)
    n, fallback, range_of_bin, range_start = plan_kernel(hist, len(keys))
    if fallback:
        return None
    return msd_sort_kernel(keys, G, n, range_of_bin, range_start)


def lsd4_fallback(keys_a, G):
    """transcription of msd_lsd4_fallback: four 8-bit passes, per-CTA ranking, counts matrix + row scan"""
    T = len(keys_a)
    ipt = (T + G * THREADS - 1) // (G * THREADS)
    assert ipt <= IPT_MAX
    chunk = THREADS * ipt
    kin, vin = np.asarray(keys_a, np.uint32).copy(), None
    for p in range(4):
        shift = 8 * p
        kout = np.zeros(T, np.uint32)
        vout = np.zeros(T, np.uint32)
        mat = np.zeros((256, G), np.int64)
        state = []
        for cta in range(G):
            cta_base = cta * chunk
            cta_valid = min(chunk, T - cta_base) if cta_base < T else 0
            tab = np.zeros((WARPS, 256), np.int64)
            items = []  # (warp, digit, rank, key, val) in (warp, item, lane) order
            for w in range(WARPS):
                warp_base = cta_base + w * (32 * ipt)
                for j in range(ipt):
                    lanes = [(l, warp_base + j * 32 + l) for l in range(32) if warp_base + j * 32 + l < T]
                    dg = {l: (int(kin[i]) >> shift) & 255 for l, i in lanes}
                    for l, i in lanes:
                        d = dg[l]
                        lower = sum(1 for m, _ in lanes if m < l and dg[m] == d)
                        items.append((w, d, tab[w, d] + lower, int(kin[i]), int(vin[i]) if vin is not None else i))
                    for d in set(dg.values()):
                        tab[w, d] += sum(1 for v in dg.values() if v == d)
            cta_count = np.zeros(256, np.int64)
            for d in range(256):
                run = 0
                for w in range(WARPS):
                    c = tab[w, d]
                    tab[w, d] = run
                    run += c
                cta_count[d] = run
                mat[d, cta] = run
            local = np.cumsum(cta_count) - cta_count          # base[256 + d]
            kv = [None] * max(cta_valid, 1)
            for w, d, rk, k, v in items:
                slot = local[d] + tab[w, d] + rk
                assert kv[slot] is None
                kv[slot] = (k, v)
            state.append((cta_valid, local, kv))
        totals = mat.sum(axis=1)
        matp = np.cumsum(mat, axis=1) - mat
        ex = np.cumsum(totals) - totals
        for cta in range(G):
            cta_valid, local, kv = state[cta]
            base = ex + matp[:, cta] - local
            for s in range(cta_valid):
                k, v = kv[s]
                dst = base[(k >> shift) & 255] + s
                kout[dst] = k
                vout[dst] = v
        kin, vin = kout, vout
    return kin, vin
