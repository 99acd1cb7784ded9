"""Tap tables / grid mappings for the tcgen05 convolution kernel (adp_tc_geom, include/adapose_b200.h)."""
from __future__ import annotations

from . import _lib as L


def _fill(g, taps):
    g.ntaps = len(taps)
    assert g.ntaps <= 28
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
