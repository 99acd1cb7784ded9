"""Tap tables / grid mappings for the tcgen05 convolution kernel (adp_tc_geom, include/adapose_b200.h)."""
from __future__ import annotations

from . import _lib as L


def _fill(g, taps):
    g.ntaps = len(taps)
    assert g.ntaps <= 32
    for i, (dz, dy, dx, wt) in enumerate(taps):
        g.dz[i], g.dy[i], g.dx[i], g.wt[i] = dz, dy, dx, wt
    return g


def strided(three_d, D, H, W, ks=3, stride=2):
    """k x k (x k) conv, padding k//2, stride 2 in every spatial axis: the M tiles walk over the OUTPUT grid and the TMA
    unit traverses the input with element stride 2 (pspnet.py:53-63 layer2.0; network_v5.py:265-272 conv1/3/5)."""
    pad = ks // 2
    kd = ks if three_d else 1
    taps = [((kz - pad) if three_d else 0, ky - pad, kx - pad, (kz * ks + ky) * ks + kx)
            for kz in range(kd) for ky in range(ks) for kx in range(ks)]
    g = _fill(L.TcGeom(), taps)
    g.in_mul, g.out_mul = stride, 1
    g.out_oz = g.out_oy = g.out_ox = 0
    oD = (D + stride - 1) // stride if three_d else 1
    oH, oW = (H + stride - 1) // stride, (W + stride - 1) // stride
    g.gD, g.gH, g.gW = oD, oH, oW
    g.oD, g.oH, g.oW = oD, oH, oW
    g.w_taps = kd * ks * ks
    return g


def transposed_classes(D, H, W):
    """ConvTranspose3d(k=3, stride=2, padding=1, output_padding=1) (network_v5.py:274-278) as 8 output-parity classes.
    o = 2 i - 1 + k: even outputs see k = 1 at i = o/2; odd outputs see k = 2 at i = (o-1)/2 and k = 0 at i = (o+1)/2.
    Each class is a stride-1 conv over the INPUT grid with 1..8 taps, written to outputs 2 i + parity."""
    axis = {0: [(1, 0)], 1: [(2, 0), (0, 1)]}     # parity -> [(k, input offset)]
    out = []
    for pz in (0, 1):
        for py in (0, 1):
            for px in (0, 1):
                taps = [(oz, oy, ox, (kz * 3 + ky) * 3 + kx)
                        for kz, oz in axis[pz] for ky, oy in axis[py] for kx, ox in axis[px]]
                g = _fill(L.TcGeom(), taps)
                g.in_mul, g.out_mul = 1, 2
                g.out_oz, g.out_oy, g.out_ox = pz, py, px
                g.gD, g.gH, g.gW = D, H, W
                g.oD, g.oH, g.oW = 2 * D, 2 * H, 2 * W
                g.w_taps = 27
                out.append(g)
    return out


def stem_s2d(S):
    """The 7x7 stride-2 pad-3 stem conv (pspnet.py:37) over a space-to-depth(2) input: output o reads pixels 2o+k-3,
    i.e. s2d cells o-2..o+1 -> a 4x4 stride-1 window over [S/2, S/2, 16] (see adp_pack_s2d)."""
    taps = [(0, ty - 2, tx - 2, ty * 4 + tx) for ty in range(4) for tx in range(4)]
    g = _fill(L.TcGeom(), taps)
    g.in_mul = g.out_mul = 1
    g.out_oz = g.out_oy = g.out_ox = 0
    g.gD, g.gH, g.gW = 1, S // 2, S // 2
    g.oD, g.oH, g.oW = 1, S // 2, S // 2
    g.w_taps = 16
    return g


def stem_s2d_split(S):
    """stem_s2d with the fp16 weight split unrolled into the tap table: every window position appears twice, reading weight slab t
    (hi) and then 16 + t (lo), so that the single-pass slab kernel (one halo slab per tile, all 32 weight slabs resident) computes
    A W_hi + A W_lo in one accumulator IN THE ORDER of the two-pass generic kernel (per K step: hi, then lo).  The generic kernel
    moves 16 x 128 rows of 32 bytes per tile and pass and is bound by the TMA unit's rows per cycle (0.40 ms per 148 frames); the
    slab is 209 rows."""
    g = stem_s2d(S)
    taps = [(0, ty - 2, tx - 2, h * 16 + ty * 4 + tx) for ty in range(4) for tx in range(4) for h in range(2)]
    _fill(g, taps)
    g.w_taps = 32
    return g


def stem_s2d_weights(w):
    """[64,3,7,7] torch weights -> [16 taps, 64, 16 ch] with ch = (py*2+px)*3 + c (zeros where the 7x7 window ends)."""
    import torch
    cout = w.shape[0]
    out = torch.zeros(16, cout, 16)
    for ty in range(4):
        for tx in range(4):
            for py in range(2):
                for px in range(2):
                    ky, kx = 2 * (ty - 2) + py + 3, 2 * (tx - 2) + px + 3
                    if 0 <= ky < 7 and 0 <= kx < 7:
                        for c in range(3):
                            out[ty * 4 + tx, :, (py * 2 + px) * 3 + c] = w[:, c, ky, kx]
    return out


def conv0_ring_weights(w):
    """[8,32,3,3,3] torch weights -> [3 kx][4 chunks of 8 channels][80 n = (kz*3+ky)*8 + co][8 ch] for adp_conv0_run (rows 72..79
    zero): the depth and row taps are folded into the MMA's N dimension (csrc/conv0_ring.cu)."""
    import torch
    out = torch.zeros(3, 4, 80, 8)
    for kx in range(3):
        for kz in range(3):
            for ky in range(3):
                blk = w[:, :, kz, ky, kx]                       # [co, c]
                n0 = (kz * 3 + ky) * 8
                out[kx, :, n0:n0 + 8, :] = blk.reshape(8, 4, 8).permute(1, 0, 2)
    return out.contiguous()


# ---- level-0 tensors of the 3-D U-Net in space-to-depth(2) layout: [B, D/2, H/2, W/2, 64], channel = (pz*4+py*2+px)*8 + c
def _s2d_grid(g, D2, H2, W2, ntaps):
    g.in_mul = g.out_mul = 1
    g.out_oz = g.out_oy = g.out_ox = 0
    g.gD, g.gH, g.gW = D2, H2, W2
    g.oD, g.oH, g.oW = D2, H2, W2
    g.w_taps = ntaps
    return g


_S2D_OFFS = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]


def strided_s2d(D2, H2, W2):
    """3x3x3 stride-2 pad-1 conv (network_v5.py:265 conv1) over an s2d(2) input: output o reads inputs 2o + k - 1, i.e.
    k = 0 -> cell o-1 parity 1, k = 1 -> cell o parity 0, k = 2 -> cell o parity 1: a 2x2x2 stride-1 window (cell offsets
    -1, 0 per axis) over 64 channels.  The M tiles walk the output grid = the s2d grid [D2,H2,W2]."""
    taps = [(a - 1, b - 1, c - 1, i) for i, (a, b, c) in enumerate(_S2D_OFFS)]
    return _s2d_grid(_fill(L.TcGeom(), taps), D2, H2, W2, 8)


def strided_s2d_weights(w, cin_real=8):
    """[Cout, Cin, 3,3,3] torch weights -> [8 taps (cell offset -1/0 per axis), Cout, 64] (zero where the window ends)."""
    import torch
    kmap = {(0, 1): 0, (1, 0): 1, (1, 1): 2}            # (tap index 0 = offset -1 / 1 = offset 0, parity) -> k
    cout = w.shape[0]
    out = torch.zeros(8, cout, 64)
    for t, offs in enumerate(_S2D_OFFS):
        for par in _S2D_OFFS:
            ks = [kmap.get((o, p)) for o, p in zip(offs, par)]
            if None in ks:
                continue
            pi = par[0] * 4 + par[1] * 2 + par[2]
            out[t, :, pi * 8:pi * 8 + cin_real] = w[:, :cin_real, ks[0], ks[1], ks[2]]
    return out.contiguous()


def transposed_s2d(D2, H2, W2):
    """ConvTranspose3d(k=3, s=2, p=1, output_padding=1) (network_v5.py:278 conv11) writing its output in s2d(2) layout:
    output 2v + p = 2i - 1 + k  ->  p = 0: (i = v, k = 1);  p = 1: (i = v, k = 2), (i = v + 1, k = 0).  A 2x2x2 stride-1
    window (cell offsets 0, +1) over the input grid [D2,H2,W2] producing 8 parities x Cout channels per cell."""
    taps = [(a, b, c, i) for i, (a, b, c) in enumerate(_S2D_OFFS)]
    return _s2d_grid(_fill(L.TcGeom(), taps), D2, H2, W2, 8)


def transposed_s2d_weights(w):
    """[Cin, Cout, 3,3,3] torch weights -> [8 taps (input offset 0/+1 per axis), 8 * Cout (parity-major), Cin]."""
    import torch
    kmap = {(0, 0): 1, (0, 1): 2, (1, 1): 0}            # (input offset, output parity) -> k
    cin, cout = w.shape[0], w.shape[1]
    out = torch.zeros(8, 8 * cout, cin)
    for t, offs in enumerate(_S2D_OFFS):
        for par in _S2D_OFFS:
            ks = [kmap.get((o, p)) for o, p in zip(offs, par)]
            if None in ks:
                continue
            pi = par[0] * 4 + par[1] * 2 + par[2]
            out[t, pi * cout:(pi + 1) * cout, :] = w[:, :, ks[0], ks[1], ks[2]].t()
    return out.contiguous()


def to_s2d(x):
    """[B, D, H, W, C] -> [B, D/2, H/2, W/2, 8*C] (channel = (pz*4+py*2+px)*C + c)."""
    B, D, H, W, Cn = x.shape
    return (x.reshape(B, D // 2, 2, H // 2, 2, W // 2, 2, Cn).permute(0, 1, 3, 5, 2, 4, 6, 7)
            .reshape(B, D // 2, H // 2, W // 2, 8 * Cn).contiguous())


def from_s2d(x):
    """inverse of :func:`to_s2d`."""
    B, D2, H2, W2, C8 = x.shape
    Cn = C8 // 8
    return (x.reshape(B, D2, H2, W2, 2, 2, 2, Cn).permute(0, 1, 4, 2, 5, 3, 6, 7)
            .reshape(B, 2 * D2, 2 * H2, 2 * W2, Cn).contiguous())
