"""ORACLE — test infrastructure only.  Second, independent restatement (numpy, vectorised
over blocks) of llama.cpp's K-quant reference quantizers, used to cross-check
oracle/ggml_quants.c byte-for-byte (SURVEY.md §7 hard part 1: "two independent GGUF
implementations must agree").  Follows SURVEY.md §D.4 (Q4_K / Q5_K, `make_qkx2_quants`; Q2_K uses the same
search with use_mad), §D.5 (Q6_K, `make_qx_quants` rmse_type=1) and llama.cpp `make_q3_quants` (Q3_K).

Every in-block accumulation is a sequential fp32 sum in element order (python loop over the
32 / 16 elements of a sub-block; numpy vectors run over blocks), so rounding matches the C
evaluation order exactly.  Parity unpinned against llama.cpp itself — see ggml_quants.c.
"""
import numpy as np

f32 = np.float32


def nearest_int(v):
    """12582912.f magic-constant rounding (round-half-even), SURVEY §D."""
    val = (v.astype(f32) + f32(12582912.0)).astype(f32)
    i = val.view(np.int32)
    return (i & 0x007FFFFF) - 0x00400000


def _qkx2(x, w, nmax, rmin, rdelta, nstep, use_mad=False):
    """x, w: [B, n] fp32.  Returns scale [B], the_min [B], L [B, n] (uint8)."""
    B, n = x.shape
    mn = x[:, 0].copy()
    mx = x[:, 0].copy()
    sum_w = w[:, 0].copy()
    sum_x = (sum_w * x[:, 0]).astype(f32)
    for i in range(1, n):
        mn = np.where(x[:, i] < mn, x[:, i], mn)
        mx = np.where(x[:, i] > mx, x[:, i], mx)
        sum_w = (sum_w + w[:, i]).astype(f32)
        sum_x = (sum_x + (w[:, i] * x[:, i]).astype(f32)).astype(f32)
    mn = np.where(mn > 0, f32(0), mn).astype(f32)
    flat = mx == mn
    rng = np.where(flat, f32(1), (mx - mn).astype(f32)).astype(f32)
    iscale = (f32(nmax) / rng).astype(f32)
    scale = (f32(1) / iscale).astype(f32)
    L = np.empty((B, n), dtype=np.int32)
    best = np.zeros(B, dtype=f32)
    for i in range(n):
        l = np.clip(nearest_int((iscale * (x[:, i] - mn).astype(f32)).astype(f32)), 0, nmax)
        L[:, i] = l
        diff = (((scale * l.astype(f32)).astype(f32) + mn).astype(f32) - x[:, i]).astype(f32)
        diff = np.abs(diff) if use_mad else (diff * diff).astype(f32)
        best = (best + (w[:, i] * diff).astype(f32)).astype(f32)
    cur_min = mn.copy()
    for step in range(nstep + 1):
        isc = ((f32(rmin) + f32(rdelta) * f32(step)).astype(f32) + f32(nmax)).astype(f32)
        with np.errstate(all="ignore"):
            iscale = (isc / np.where(flat, f32(1), (mx - cur_min).astype(f32))).astype(f32)
        Laux = np.empty((B, n), dtype=np.int32)
        sum_l = np.zeros(B, dtype=f32)
        sum_l2 = np.zeros(B, dtype=f32)
        sum_xl = np.zeros(B, dtype=f32)
        for i in range(n):
            # `min` is overwritten on adoption (min = this_min), so both the grid origin and
            # the (max - min) span follow the best candidate so far.
            l = np.clip(nearest_int((iscale * (x[:, i] - cur_min).astype(f32)).astype(f32)), 0, nmax)
            Laux[:, i] = l
            lf = l.astype(f32)
            wl = (w[:, i] * lf).astype(f32)
            sum_l = (sum_l + wl).astype(f32)
            sum_l2 = (sum_l2 + (wl * lf).astype(f32)).astype(f32)
            sum_xl = (sum_xl + (wl * x[:, i]).astype(f32)).astype(f32)
        D = ((sum_w * sum_l2).astype(f32) - (sum_l * sum_l).astype(f32)).astype(f32)
        ok = D > 0
        Ds = np.where(ok, D, f32(1)).astype(f32)
        with np.errstate(all="ignore"):
            this_scale = (((sum_w * sum_xl).astype(f32) - (sum_x * sum_l).astype(f32)).astype(f32) / Ds).astype(f32)
            this_min = (((sum_l2 * sum_x).astype(f32) - (sum_l * sum_xl).astype(f32)).astype(f32) / Ds).astype(f32)
            pos = this_min > 0
            alt = (sum_xl / np.where(sum_l2 == 0, f32(1), sum_l2)).astype(f32)
        this_scale = np.where(pos, alt, this_scale).astype(f32)
        this_min = np.where(pos, f32(0), this_min).astype(f32)
        err = np.zeros(B, dtype=f32)
        for i in range(n):
            diff = (((this_scale * Laux[:, i].astype(f32)).astype(f32) + this_min).astype(f32) - x[:, i]).astype(f32)
            diff = np.abs(diff) if use_mad else (diff * diff).astype(f32)
            err = (err + (w[:, i] * diff).astype(f32)).astype(f32)
        adopt = ok & (err < best) & ~flat
        L[adopt] = Laux[adopt]
        best = np.where(adopt, err, best).astype(f32)
        scale = np.where(adopt, this_scale, scale).astype(f32)
        cur_min = np.where(adopt, this_min, cur_min).astype(f32)
    scale = np.where(flat, f32(0), scale).astype(f32)
    L[flat] = 0
    the_min = np.where(flat, -mn, -cur_min).astype(f32)
    return scale, the_min, L


def _k45_front(xb, nmax, rmin, rdelta, nstep):
    """xb: [B, 256].  Returns packed scales [B,12] u8, d,dmin (fp16 as u16), L [B,256]."""
    B = xb.shape[0]
    xs = xb.reshape(B * 8, 32)
    sum_x2 = np.zeros(B * 8, dtype=f32)
    for l in range(32):
        sum_x2 = (sum_x2 + (xs[:, l] * xs[:, l]).astype(f32)).astype(f32)
    av_x = np.sqrt((sum_x2 / f32(32)).astype(f32)).astype(f32)
    w = (av_x[:, None] + np.abs(xs)).astype(f32)
    sc, mins, L = _qkx2(xs, w, nmax, rmin, rdelta, nstep)
    sc = sc.reshape(B, 8)
    mins = mins.reshape(B, 8)
    L = L.reshape(B, 256)
    max_scale = np.zeros(B, dtype=f32)
    max_min = np.zeros(B, dtype=f32)
    for j in range(8):
        max_scale = np.where(sc[:, j] > max_scale, sc[:, j], max_scale)
        max_min = np.where(mins[:, j] > max_min, mins[:, j], max_min)
    with np.errstate(all="ignore"):
        inv_scale = np.where(max_scale > 0, f32(63) / np.where(max_scale > 0, max_scale, f32(1)), f32(0)).astype(f32)
        inv_min = np.where(max_min > 0, f32(63) / np.where(max_min > 0, max_min, f32(1)), f32(0)).astype(f32)
    scales = np.zeros((B, 12), dtype=np.uint8)
    ls_all = np.empty((B, 8), dtype=np.int32)
    lm_all = np.empty((B, 8), dtype=np.int32)
    for j in range(8):
        ls = np.minimum(63, (nearest_int((inv_scale * sc[:, j]).astype(f32)) & 0xFF))
        lm = np.minimum(63, (nearest_int((inv_min * mins[:, j]).astype(f32)) & 0xFF))
        ls_all[:, j], lm_all[:, j] = ls, lm
        if j < 4:
            scales[:, j] = ls
            scales[:, j + 4] = lm
        else:
            scales[:, j + 4] = (ls & 0xF) | ((lm & 0xF) << 4)
            scales[:, j - 4] |= ((ls >> 4) << 6).astype(np.uint8)
            scales[:, j] |= ((lm >> 4) << 6).astype(np.uint8)
    d16 = (max_scale / f32(63)).astype(f32).astype(np.float16)
    m16 = (max_min / f32(63)).astype(f32).astype(np.float16)
    dq = d16.astype(f32)
    mq = m16.astype(f32)
    for j in range(8):
        d = (dq * ls_all[:, j].astype(f32)).astype(f32)
        dm = (mq * lm_all[:, j].astype(f32)).astype(f32)
        nz = d != 0
        ds = np.where(nz, d, f32(1)).astype(f32)
        for ii in range(32):
            l = np.clip(nearest_int(((xb[:, 32 * j + ii] + dm).astype(f32) / ds).astype(f32)), 0, nmax)
            L[:, 32 * j + ii] = np.where(nz, l, L[:, 32 * j + ii])
    return scales, d16.view(np.uint16), m16.view(np.uint16), L


def quantize_q4_K(x):
    x = np.ascontiguousarray(x, dtype=f32)
    nrows, ncols = x.shape
    xb = x.reshape(-1, 256)
    B = xb.shape[0]
    scales, d, dmin, L = _k45_front(xb, 15, -1.0, 0.1, 20)
    out = np.zeros((B, 144), dtype=np.uint8)
    out[:, 0:2] = d.view(np.uint8).reshape(B, 2)
    out[:, 2:4] = dmin.view(np.uint8).reshape(B, 2)
    out[:, 4:16] = scales
    Lc = L.reshape(B, 4, 2, 32)
    out[:, 16:144] = (Lc[:, :, 0, :] | (Lc[:, :, 1, :] << 4)).astype(np.uint8).reshape(B, 128)
    return out.reshape(nrows, -1)


def quantize_q5_K(x):
    x = np.ascontiguousarray(x, dtype=f32)
    nrows, ncols = x.shape
    xb = x.reshape(-1, 256)
    B = xb.shape[0]
    scales, d, dmin, L = _k45_front(xb, 31, -0.5, 0.1, 15)
    out = np.zeros((B, 176), dtype=np.uint8)
    out[:, 0:2] = d.view(np.uint8).reshape(B, 2)
    out[:, 2:4] = dmin.view(np.uint8).reshape(B, 2)
    out[:, 4:16] = scales
    Lc = L.reshape(B, 4, 2, 32)
    qh = np.zeros((B, 32), dtype=np.int32)
    for c in range(4):
        qh |= ((Lc[:, c, 0, :] > 15).astype(np.int32) << (2 * c))
        qh |= ((Lc[:, c, 1, :] > 15).astype(np.int32) << (2 * c + 1))
    out[:, 16:48] = qh.astype(np.uint8)
    lo = Lc & 0xF
    out[:, 48:176] = (lo[:, :, 0, :] | (lo[:, :, 1, :] << 4)).astype(np.uint8).reshape(B, 128)
    return out.reshape(nrows, -1)


def _qx_rmse1(x, nmax):
    """x: [B, n].  Returns scale [B], L [B, n] (already offset by +nmax)."""
    B, n = x.shape
    amax = np.zeros(B, dtype=f32)
    mx = np.zeros(B, dtype=f32)
    for i in range(n):
        ax = np.abs(x[:, i])
        upd = ax > amax
        amax = np.where(upd, ax, amax)
        mx = np.where(upd, x[:, i], mx)
    tiny = amax < f32(1e-15)
    mxs = np.where(tiny, f32(1), mx).astype(f32)

    def sums(iscale):
        sumlx = np.zeros(B, dtype=f32)
        suml2 = np.zeros(B, dtype=f32)
        Ls = np.empty((B, n), dtype=np.int32)
        for i in range(n):
            l = np.clip(nearest_int((iscale * x[:, i]).astype(f32)), -nmax, nmax - 1)
            Ls[:, i] = l
            lf = l.astype(f32)
            w = (x[:, i] * x[:, i]).astype(f32)
            sumlx = (sumlx + ((w * x[:, i]).astype(f32) * lf).astype(f32)).astype(f32)
            suml2 = (suml2 + ((w * lf).astype(f32) * lf).astype(f32)).astype(f32)
        return sumlx, suml2, Ls

    iscale = (f32(-nmax) / mxs).astype(f32)
    sumlx, suml2, L = sums(iscale)
    L = L + nmax
    with np.errstate(all="ignore"):
        scale = np.where(suml2 != 0, sumlx / np.where(suml2 != 0, suml2, f32(1)), f32(0)).astype(f32)
    best = (scale * sumlx).astype(f32)
    for s in range(-9, 10):
        if s == 0:
            continue
        iscale = (-(f32(nmax) + (f32(0.1) * f32(s)).astype(f32)).astype(f32) / mxs).astype(f32)
        sumlx, suml2, Ls = sums(iscale)
        adopt = (suml2 > 0) & ((sumlx * sumlx).astype(f32) > (best * suml2).astype(f32))
        L[adopt] = Ls[adopt] + nmax
        with np.errstate(all="ignore"):
            ns = (sumlx / np.where(suml2 != 0, suml2, f32(1))).astype(f32)
        scale = np.where(adopt, ns, scale).astype(f32)
        best = np.where(adopt, (ns * sumlx).astype(f32), best).astype(f32)
    scale = np.where(tiny, f32(0), scale).astype(f32)
    L[tiny] = 0
    return scale, L


def quantize_q6_K(x):
    x = np.ascontiguousarray(x, dtype=f32)
    nrows, ncols = x.shape
    xb = x.reshape(-1, 256)
    B = xb.shape[0]
    sc, L = _qx_rmse1(xb.reshape(B * 16, 16), 32)
    sc = sc.reshape(B, 16)
    L = L.reshape(B, 256)
    max_scale = np.zeros(B, dtype=f32)
    max_abs = np.zeros(B, dtype=f32)
    for ib in range(16):
        a = np.abs(sc[:, ib])
        upd = a > max_abs
        max_abs = np.where(upd, a, max_abs)
        max_scale = np.where(upd, sc[:, ib], max_scale)
    zero = max_abs < f32(1e-15)
    ms = np.where(zero, f32(1), max_scale).astype(f32)
    iscale = (f32(-128.0) / ms).astype(f32)
    d16 = (f32(1) / iscale).astype(f32).astype(np.float16)
    dq = d16.astype(f32)
    q_sc = np.empty((B, 16), dtype=np.int32)
    for ib in range(16):
        q_sc[:, ib] = np.minimum(127, nearest_int((iscale * sc[:, ib]).astype(f32)))
    q_sc8 = q_sc.astype(np.int8)
    for j in range(16):
        d = (dq * q_sc8[:, j].astype(f32)).astype(f32)
        nz = d != 0
        ds = np.where(nz, d, f32(1)).astype(f32)
        for ii in range(16):
            l = np.clip(nearest_int((xb[:, 16 * j + ii] / ds).astype(f32)), -32, 31) + 32
            L[:, 16 * j + ii] = np.where(nz, l, L[:, 16 * j + ii])
    out = np.zeros((B, 210), dtype=np.uint8)
    Lh = L.reshape(B, 2, 4, 32)
    lo = Lh & 0xF
    hi = Lh >> 4
    ql = np.empty((B, 2, 64), dtype=np.int32)
    ql[:, :, 0:32] = lo[:, :, 0, :] | (lo[:, :, 2, :] << 4)
    ql[:, :, 32:64] = lo[:, :, 1, :] | (lo[:, :, 3, :] << 4)
    qh = hi[:, :, 0, :] | (hi[:, :, 1, :] << 2) | (hi[:, :, 2, :] << 4) | (hi[:, :, 3, :] << 6)
    out[:, 0:128] = ql.reshape(B, 128).astype(np.uint8)
    out[:, 128:192] = qh.reshape(B, 64).astype(np.uint8)
    out[:, 192:208] = q_sc8.view(np.uint8)
    out[:, 208:210] = d16.view(np.uint8).reshape(B, 2)
    out[zero] = 0
    return out.reshape(nrows, -1)


def _pack2(L):
    """[B,256] 2-bit codes -> [B,64]: byte l of each 128-half = L[l] | L[l+32]<<2 | L[l+64]<<4 | L[l+96]<<6."""
    B = L.shape[0]
    Lh = L.reshape(B, 2, 4, 32)
    return (Lh[:, :, 0, :] | (Lh[:, :, 1, :] << 2) | (Lh[:, :, 2, :] << 4) | (Lh[:, :, 3, :] << 6)).astype(np.uint8).reshape(B, 64)


def quantize_q2_K(x):
    """llama.cpp quantize_row_q2_K_ref: qkx2(16, 3, |x| weights, -0.5, 0.1, 15, use_mad), 4-bit scales/mins."""
    x = np.ascontiguousarray(x, dtype=f32)
    nrows, ncols = x.shape
    xb = x.reshape(-1, 256)
    B = xb.shape[0]
    xs = xb.reshape(B * 16, 16)
    sc, mins, L = _qkx2(xs, np.abs(xs).astype(f32), 3, -0.5, 0.1, 15, use_mad=True)
    sc, mins, L = sc.reshape(B, 16), mins.reshape(B, 16), L.reshape(B, 256)
    max_scale = np.zeros(B, dtype=f32)
    max_min = np.zeros(B, dtype=f32)
    for j in range(16):
        max_scale = np.where(sc[:, j] > max_scale, sc[:, j], max_scale)
        max_min = np.where(mins[:, j] > max_min, mins[:, j], max_min)
    has_s, has_m = max_scale > 0, max_min > 0
    with np.errstate(all="ignore"):
        isc = (f32(15) / np.where(has_s, max_scale, f32(1))).astype(f32)
        imn = (f32(15) / np.where(has_m, max_min, f32(1))).astype(f32)
    ls = np.where(has_s[:, None], nearest_int((isc[:, None] * sc).astype(f32)), 0)
    lm = np.where(has_m[:, None], nearest_int((imn[:, None] * mins).astype(f32)), 0)
    scales = ((ls & 0xFF) | ((lm << 4) & 0xFF)).astype(np.uint8)
    d16 = np.where(has_s, (max_scale / f32(15)).astype(f32), f32(0)).astype(np.float16)
    m16 = np.where(has_m, (max_min / f32(15)).astype(f32), f32(0)).astype(np.float16)
    dq, mq = d16.astype(f32), m16.astype(f32)
    for j in range(16):
        d = (dq * (scales[:, j] & 0xF).astype(f32)).astype(f32)
        dm = (mq * (scales[:, j] >> 4).astype(f32)).astype(f32)
        nz = d != 0
        ds = np.where(nz, d, f32(1)).astype(f32)
        for ii in range(16):
            l = np.clip(nearest_int(((xb[:, 16 * j + ii] + dm).astype(f32) / ds).astype(f32)), 0, 3)
            L[:, 16 * j + ii] = np.where(nz, l, L[:, 16 * j + ii])
    out = np.zeros((B, 84), dtype=np.uint8)
    out[:, 0:16] = scales
    out[:, 16:80] = _pack2(L)
    out[:, 80:82] = d16.view(np.uint8).reshape(B, 2)
    out[:, 82:84] = m16.view(np.uint8).reshape(B, 2)
    return out.reshape(nrows, -1)


def _q3_rmse(x, nmax):
    """llama.cpp make_q3_quants(do_rmse=true).  x: [B, n] -> scale [B], L [B, n] offset by +nmax."""
    B, n = x.shape
    amax = np.zeros(B, dtype=f32)
    mx = np.zeros(B, dtype=f32)
    for i in range(n):
        ax = np.abs(x[:, i])
        upd = ax > amax
        amax = np.where(upd, ax, amax)
        mx = np.where(upd, x[:, i], mx)
    tiny = amax < f32(1e-15)
    iscale = (f32(-nmax) / np.where(tiny, f32(1), mx)).astype(f32)
    L = np.empty((B, n), dtype=np.int32)
    sumlx = np.zeros(B, dtype=f32)
    suml2 = np.zeros(B, dtype=f32)
    w = (x * x).astype(f32)
    wx = (w * x).astype(f32)
    for i in range(n):
        l = np.clip(nearest_int((iscale * x[:, i]).astype(f32)), -nmax, nmax - 1)
        L[:, i] = l
        lf = l.astype(f32)
        sumlx = (sumlx + (wx[:, i] * lf).astype(f32)).astype(f32)
        suml2 = (suml2 + ((w[:, i] * lf).astype(f32) * lf).astype(f32)).astype(f32)
    active = ~tiny
    with np.errstate(all="ignore"):
        for _ in range(5):
            changed = np.zeros(B, dtype=bool)
            for i in range(n):
                lf = L[:, i].astype(f32)
                slx = (sumlx - (wx[:, i] * lf).astype(f32)).astype(f32)
                sl2 = (suml2 - ((w[:, i] * lf).astype(f32) * lf).astype(f32)).astype(f32)
                pos = slx > 0
                q = ((x[:, i] * sl2).astype(f32) / np.where(pos, slx, f32(1))).astype(f32)
                new_l = np.clip(nearest_int(q), -nmax, nmax - 1)
                nf = new_l.astype(f32)
                slx2 = (slx + (wx[:, i] * nf).astype(f32)).astype(f32)
                sl22 = (sl2 + ((w[:, i] * nf).astype(f32) * nf).astype(f32)).astype(f32)
                better = ((slx2 * slx2).astype(f32) * suml2).astype(f32) > ((sumlx * sumlx).astype(f32) * sl22).astype(f32)
                adopt = active & pos & (new_l != L[:, i]) & (sl22 > 0) & better
                L[:, i] = np.where(adopt, new_l, L[:, i])
                sumlx = np.where(adopt, slx2, sumlx).astype(f32)
                suml2 = np.where(adopt, sl22, suml2).astype(f32)
                changed |= adopt
            active = active & changed     # a block that made no change stops sweeping
            if not active.any():
                break
        scale = (sumlx / suml2).astype(f32)
    scale = np.where(tiny, f32(0), scale).astype(f32)
    L = L + nmax
    L[tiny] = 0
    return scale, L


def quantize_q3_K(x):
    x = np.ascontiguousarray(x, dtype=f32)
    nrows, ncols = x.shape
    xb = x.reshape(-1, 256)
    B = xb.shape[0]
    sc, L = _q3_rmse(xb.reshape(B * 16, 16), 4)
    sc, L = sc.reshape(B, 16), L.reshape(B, 256)
    max_scale = np.zeros(B, dtype=f32)
    amax = np.zeros(B, dtype=f32)
    for j in range(16):
        a = np.abs(sc[:, j])
        upd = a > amax
        amax = np.where(upd, a, amax)
        max_scale = np.where(upd, sc[:, j], max_scale)
    has = max_scale != 0
    with np.errstate(all="ignore"):
        iscale = (f32(-32) / np.where(has, max_scale, f32(1))).astype(f32)
        d16 = np.where(has, (f32(1) / iscale).astype(f32), f32(0)).astype(np.float16)
    l6 = np.clip(nearest_int((iscale[:, None] * sc).astype(f32)).astype(np.int8).astype(np.int32), -32, 31) + 32
    l6 = np.where(has[:, None], l6, 0)
    scales = np.zeros((B, 12), dtype=np.int32)
    for j in range(16):
        if j < 8:
            scales[:, j] = l6[:, j] & 0xF
        else:
            scales[:, j - 8] |= (l6[:, j] & 0xF) << 4
        scales[:, j % 4 + 8] |= (l6[:, j] >> 4) << (2 * (j // 4))
    dq = d16.astype(f32)
    sc6 = np.where(has[:, None], l6 - 32, -32)     # decoded exactly as the C re-reads the packed bytes
    for j in range(16):
        d = (dq * sc6[:, j].astype(f32)).astype(f32)
        nz = d != 0
        ds = np.where(nz, d, f32(1)).astype(f32)
        for ii in range(16):
            l = np.clip(nearest_int((xb[:, 16 * j + ii] / ds).astype(f32)), -4, 3) + 4
            L[:, 16 * j + ii] = np.where(nz, l, L[:, 16 * j + ii])
    hi = (L > 3).astype(np.int32).reshape(B, 8, 32)
    hmask = np.zeros((B, 32), dtype=np.int32)
    for b in range(8):
        hmask |= hi[:, b, :] << b
    out = np.zeros((B, 110), dtype=np.uint8)
    out[:, 0:32] = hmask.astype(np.uint8)
    out[:, 32:96] = _pack2(L & 3)
    out[:, 96:108] = scales.astype(np.uint8)
    out[:, 108:110] = d16.view(np.uint8).reshape(B, 2)
    return out.reshape(nrows, -1)


KVALUES_IQ4NL = np.array([-127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113], dtype=f32)


def _best_index(x):
    """llama.cpp best_index_int8 over KVALUES_IQ4NL: nearest table value, ties to the upper one."""
    v = KVALUES_IQ4NL
    mu = np.clip(np.searchsorted(v, x, side="right"), 1, 15)      # first index with v[mu] > x (binary search result)
    lower = (x - v[mu - 1]).astype(f32) < (v[mu] - x).astype(f32)
    idx = np.where(lower, mu - 1, mu)
    idx = np.where(x <= v[0], 0, idx)
    return np.where(x >= v[15], 15, idx)


def quantize_iq4_nl(x):
    """quantize_row_iq4_nl_impl(32, 32, quant_weights=NULL, ntry=7), vectorised over blocks."""
    x = np.ascontiguousarray(x, dtype=f32)
    nrows, ncols = x.shape
    xb = x.reshape(-1, 32)
    B = xb.shape[0]
    w = (xb * xb).astype(f32)
    amax = np.zeros(B, dtype=f32)
    mx = np.zeros(B, dtype=f32)
    for j in range(32):
        ax = np.abs(xb[:, j])
        upd = ax > amax
        amax = np.where(upd, ax, amax)
        mx = np.where(upd, xb[:, j], mx)
    tiny = amax < f32(1e-15)
    mxs = np.where(tiny, f32(1), mx).astype(f32)

    def sums(idv):
        sumqx = np.zeros(B, dtype=f32)
        sumq2 = np.zeros(B, dtype=f32)
        for j in range(32):
            q = KVALUES_IQ4NL[_best_index((idv * xb[:, j]).astype(f32))]
            wq = (w[:, j] * q).astype(f32)
            sumqx = (sumqx + (wq * xb[:, j]).astype(f32)).astype(f32)
            sumq2 = (sumq2 + (wq * q).astype(f32)).astype(f32)
        return sumqx, sumq2

    with np.errstate(all="ignore"):
        d = (-mxs / f32(-127)).astype(f32)
        sumqx, sumq2 = sums((f32(1) / d).astype(f32))
        d = (sumqx / sumq2).astype(f32)
        best = (d * sumqx).astype(f32)
        for itry in range(-7, 8):
            idv = (f32(itry - 127) / mxs).astype(f32)
            sumqx, sumq2 = sums(idv)
            adopt = (sumq2 > 0) & ((sumqx * sumqx).astype(f32) > (best * sumq2).astype(f32))
            nd = (sumqx / np.where(sumq2 != 0, sumq2, f32(1))).astype(f32)
            d = np.where(adopt, nd, d).astype(f32)
            best = np.where(adopt, (nd * sumqx).astype(f32), best).astype(f32)
        scale = np.where(tiny, f32(0), d).astype(f32)
        idv = np.where(scale != 0, f32(1) / np.where(scale != 0, scale, f32(1)), f32(0)).astype(f32)
    L = np.empty((B, 32), dtype=np.int32)
    for j in range(32):
        L[:, j] = _best_index((idv * xb[:, j]).astype(f32))
    out = np.zeros((B, 18), dtype=np.uint8)
    out[:, 0:2] = scale.astype(np.float16).view(np.uint8).reshape(B, 2)
    out[:, 2:18] = (L[:, :16] | (L[:, 16:] << 4)).astype(np.uint8)
    return out.reshape(nrows, -1)


QUANTIZE = {"IQ4_NL": quantize_iq4_nl, "Q2_K": quantize_q2_K, "Q3_K": quantize_q3_K, "Q4_K": quantize_q4_K, "Q5_K": quantize_q5_K, "Q6_K": quantize_q6_K}
