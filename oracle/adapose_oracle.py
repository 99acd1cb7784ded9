"""CPU oracle for the AdaPose hot path -- TEST INFRASTRUCTURE, NOT THE PRODUCT.

A from-scratch restatement (torch-CPU fp32 + numpy fp64) of what hyperplane-lab/RGBManip executes for
``AdaPoseEstimator_v5.estimate`` in eval mode.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the product path
(``rgbmanip_b200``) never does and fails loudly without its CUDA library.

Parity pin: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the *reference itself*, imported from /root/reference by ``oracle/make_golden.py`` in
the build container; the resulting vectors live in ``tests/golden/`` and ``tests/test_oracle_golden.py``
checks this file against them.

Every function cites the reference lines it restates (paths relative to the reference root;
``ADA = models/pose_estimator/AdaPose``).
"""
from __future__ import annotations

import math
import warnings

import numpy as np
import torch
import torch.nn.functional as F

IMG_SIZE = 224
N_PTS = 1024
N_DEPTH = 24
DEPTH_MIN = 0.1
DEPTH_INTERVAL = 0.1
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
BN_EPS = 1e-5

DEFAULT_BBOX = np.asarray([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1],
                           [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], dtype=np.float64) + 10.0
"""Failure sentinel (ADA/interface_v5.py:232-241)."""


# ----------------------------------------------------------------------------------------------
# host-side preprocessing  (ADA/interface_v5.py:58-170, ADA/lib/utils.py:10-38)
# ----------------------------------------------------------------------------------------------
def get_bbox(y1, x1, y2, x2, img_h=480, img_w=640):
    """Square crop window, multiple of 40 and <= 440, shifted inside the frame (utils.py:10-38)."""
    ws = (max(y2 - y1, x2 - x1) // 40 + 1) * 40
    ws = min(ws, 440)
    cy, cx = (y1 + y2) // 2, (x1 + x2) // 2
    rmin, rmax = cy - ws // 2, cy + ws // 2
    cmin, cmax = cx - ws // 2, cx + ws // 2
    if rmin < 0:
        rmax -= rmin
        rmin = 0
    if cmin < 0:
        cmax -= cmin
        cmin = 0
    if rmax > img_h:
        rmin -= rmax - img_h
        rmax = img_h
    if cmax > img_w:
        cmin -= cmax - img_w
        cmax = img_w
    return int(rmin), int(rmax), int(cmin), int(cmax)


def nearest_src_index(dst_size, src_size):
    """Source index table of ``cv2.resize(..., INTER_NEAREST)`` (OpenCV resizeNN: floor(x * (1/fx)) with
    fx = dst/src evaluated in double, clamped to src-1).  Used at interface_v5.py:123."""
    inv = 1.0 / (float(dst_size) / float(src_size))
    idx = np.floor(np.arange(dst_size, dtype=np.float64) * inv).astype(np.int64)
    return np.minimum(idx, src_size - 1)


def resize_nearest(a, dst_size):
    iy = nearest_src_index(dst_size, a.shape[0])
    ix = nearest_src_index(dst_size, a.shape[1])
    return a[iy][:, ix]


def linear_taps(dst_size, src_size, dtype):
    """(i0, i1, w1) of ``cv2.resize(..., INTER_LINEAR)``: half-pixel centres, no antialias, edge clamp
    (OpenCV resize.cpp linear coefficient table).  Used at interface_v5.py:148."""
    scale = float(src_size) / float(dst_size)
    f = (np.arange(dst_size, dtype=np.float64) + 0.5) * scale - 0.5
    i0 = np.floor(f).astype(np.int64)
    w1 = f - i0
    lo = i0 < 0
    i0[lo] = 0
    w1[lo] = 0.0
    hi = i0 >= src_size - 1
    i0[hi] = src_size - 1
    w1[hi] = 0.0
    i1 = np.minimum(i0 + 1, src_size - 1)
    return i0, i1, w1.astype(dtype)


def resize_linear(a, dst_size):
    """Bilinear resize of an HxWxC float array, computed in the array's own float type like OpenCV."""
    dt = a.dtype if a.dtype in (np.float32, np.float64) else np.float32
    a = a.astype(dt, copy=False)
    y0, y1, wy = linear_taps(dst_size, a.shape[0], dt)
    x0, x1, wx = linear_taps(dst_size, a.shape[1], dt)
    # horizontal pass first, then vertical (OpenCV HResizeLinear -> VResizeLinear)
    rows = a[:, x0] * (1 - wx)[None, :, None] + a[:, x1] * wx[None, :, None]
    return rows[y0] * (1 - wy)[:, None, None] + rows[y1] * wy[:, None, None]


def sample_choose(flat_nonzero, rng=np.random):
    """Exactly N_PTS foreground indices (interface_v5.py:124-134): random subset kept in ascending order
    when there are more (global numpy RNG, ``shuffle`` of a 0/1 selector), wrap-padding when fewer."""
    n = len(flat_nonzero)
    if n > N_PTS:
        c_mask = np.zeros(n, dtype=int)
        c_mask[:N_PTS] = 1
        rng.shuffle(c_mask)
        return flat_nonzero[c_mask.nonzero()]
    if n == 0:
        return None
    return np.pad(flat_nonzero, (0, N_PTS - n), "wrap")


def prepare_model_input(rgb, mask, intrinsic, resize_size=IMG_SIZE, rng=np.random):
    """One view: crop window, resized+normalised crop, sampled pixel indices, 2-D points, cropped K.
    Restates interface_v5.py:58-170.  Returns 4 x None for an empty mask."""
    ys, xs = np.nonzero(mask)
    if len(ys) == 0:
        return None, None, None, None
    rmin, rmax, cmin, cmax = get_bbox(int(ys.min()), int(xs.min()), int(ys.max()), int(xs.max()),
                                      rgb.shape[0], rgb.shape[1])
    resize_mask = resize_nearest(np.asarray(mask[rmin:rmax, cmin:cmax]).astype(np.float32), resize_size)
    choose = sample_choose(resize_mask.flatten().nonzero()[0], rng)
    if choose is None:
        return None, None, None, None
    crop_w = rmax - rmin
    ratio = resize_size / crop_w
    xm = (choose % resize_size).astype(np.float32)[:, None]
    ym = (choose // resize_size).astype(np.float32)[:, None]
    # the window bounds are numpy integers in the reference, so ``ratio`` is a np.float64 scalar: under NumPy >= 2 promotion
    # (this image, where the goldens were made) float32 / float64-scalar is float64; NumPy 1.x kept float32 (differs by < 2e-5 px)
    pts2d = np.concatenate((xm.astype(np.float64) / ratio + cmin, ym.astype(np.float64) / ratio + rmin), axis=-1)
    crop = resize_linear(np.asarray(rgb[rmin:rmax, cmin:cmax, :]), resize_size)
    # transforms.ToTensor on a float ndarray = HWC->CHW without rescaling; Normalize (interface_v5.py:52-54)
    mean = np.asarray(IMAGENET_MEAN, crop.dtype)[:, None, None]
    std = np.asarray(IMAGENET_STD, crop.dtype)[:, None, None]
    view = (np.transpose(crop, (2, 0, 1)) - mean) / std
    fx, fy, cx, cy = intrinsic[0, 0], intrinsic[1, 1], intrinsic[0, 2], intrinsic[1, 2]
    ccx, ccy = float(cmin + cmax) / 2, float(rmin + rmax) / 2
    csx, csy = float(cmax - cmin + 1), float(rmax - rmin + 1)
    Kp = np.eye(3)
    Kp[0, 0] = fx * ratio
    Kp[1, 1] = fy * ratio
    Kp[0, 2] = (cx - (ccx - csx / 2)) * ratio
    Kp[1, 2] = (cy - (ccy - csy / 2)) * ratio
    return view, choose, pts2d, Kp


def depth_hypotheses():
    """24 planes 0.1 ... 2.4 m (interface_v5.py:272-277)."""
    return np.arange(DEPTH_MIN, DEPTH_INTERVAL * (N_DEPTH - 0.5) + DEPTH_MIN, DEPTH_INTERVAL, dtype=np.float32)


def projection(Kp, E):
    """P = [K' E[:3,:]; 0 0 0 1] (interface_v5.py:264-267)."""
    P = np.eye(4)
    P[:3, :] = Kp @ E[:3, :]
    return P


# ----------------------------------------------------------------------------------------------
# network  (ADA/lib/pspnet.py, ADA/lib/network_v5.py) -- functional, weights = reference state_dict
# ----------------------------------------------------------------------------------------------
class Taps:
    """Optional recorder / rounding hook: ``q(name, tensor) -> tensor`` is applied to every activation
    the device pipeline materialises, so tests can (i) grab stage outputs and (ii) emulate the device's
    storage rounding when bounding its error budget.  Identity by default."""

    def __init__(self, q=None, record=None):
        self.q = q
        self.record = record
        self.store = {}

    def __call__(self, name, x):
        if self.q is not None:
            x = self.q(name, x)
        if self.record is not None and (self.record is True or name in self.record):
            self.store[name] = x.detach().clone()
        return x


_NOTAP = Taps()


def _w(sd, name):
    v = sd[name]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))


def basic_block(sd, pre, x, stride, dilation, tap, has_down):
    """relu(conv2(relu(conv1(x))) + shortcut(x)), no BatchNorm (pspnet.py:11-30)."""
    out = F.conv2d(x, _w(sd, f"{pre}.conv1.weight"), stride=stride, padding=dilation, dilation=dilation)
    out = tap(f"{pre}.conv1", F.relu(out))
    out = F.conv2d(out, _w(sd, f"{pre}.conv2.weight"), stride=1, padding=dilation, dilation=dilation)
    res = x
    if has_down:
        res = tap(f"{pre}.down", F.conv2d(x, _w(sd, f"{pre}.downsample.0.weight"), stride=stride))
    return tap(f"{pre}", F.relu(out + res))


def resnet34_dilated(sd, x, tap=_NOTAP):
    """pspnet.py:33-73: 7x7/2 conv, ReLU, 3x3/2 max-pool, layers (3,4,6,3); layer3/4 keep stride 1 and
    use dilation 2/4 in every block but their first (``_make_layer`` passes dilation only to blocks>=1)."""
    p = "img_extractor.feats"
    x = tap("conv1", F.relu(F.conv2d(x, _w(sd, f"{p}.conv1.weight"), stride=2, padding=3)))
    x = tap("maxpool", F.max_pool2d(x, kernel_size=3, stride=2, padding=1))
    for li, (blocks, stride, dil) in enumerate(((3, 1, 1), (4, 2, 1), (6, 1, 2), (3, 1, 4)), start=1):
        for b in range(blocks):
            pre = f"{p}.layer{li}.{b}"
            x = basic_block(sd, pre, x, stride if b == 0 else 1, 1 if b == 0 else dil, tap,
                            has_down=f"{pre}.downsample.0.weight" in sd)
    return x


def psp_module(sd, f, tap=_NOTAP):
    """pspnet.py:76-94: adaptive-avg-pool to b x b, 1x1 conv (no bias), ReLU, bilinear(align_corners=True)
    back to the map size, concat with the input."""
    h, w = f.shape[2:]
    priors = [f]
    for s, b in enumerate((1, 2, 3, 6)):
        y = F.relu(F.conv2d(F.adaptive_avg_pool2d(f, b), _w(sd, f"img_extractor.psp.stages.{s}.1.weight")))
        priors.append(F.interpolate(y, size=(h, w), mode="bilinear", align_corners=True))
    return tap("psp", torch.cat(priors, 1))


def psp_upsample(sd, name, x, tap=_NOTAP):
    """pspnet.py:97-107: bilinear x2 (align_corners=True) -> 3x3 conv + bias -> PReLU (single slope)."""
    x = tap(f"{name}.up", F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True))
    x = F.conv2d(x, _w(sd, f"img_extractor.{name}.conv.0.weight"), _w(sd, f"img_extractor.{name}.conv.0.bias"),
                 padding=1)
    return tap(name, F.prelu(x, _w(sd, f"img_extractor.{name}.conv.1.weight")))


def pspnet(sd, x, tap=_NOTAP):
    """pspnet.py:142-158 in eval mode (Dropout2d is the identity)."""
    f = resnet34_dilated(sd, x, tap)
    p = psp_module(sd, f, tap)
    p = psp_upsample(sd, "up_1", p, tap)
    p = psp_upsample(sd, "up_2", p, tap)
    p = psp_upsample(sd, "up_3", p, tap)
    return tap("feat", F.conv2d(p, _w(sd, "img_extractor.final.weight"), _w(sd, "img_extractor.final.bias")))


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    """Plane-sweep warp of the source feature map into the reference view (network_v5.py:378-416).
    Coordinates are normalised with (W-1)/2 but sampled with align_corners=False (torch default), zeros
    padding -- reproduced as written."""
    B, C, H, W = src_fea.shape
    D = depth_values.shape[1]
    proj = torch.matmul(src_proj, torch.inverse(ref_proj))
    rot, trans = proj[:, :3, :3], proj[:, :3, 3:4]
    y, x = torch.meshgrid(torch.arange(0, H, dtype=torch.float32), torch.arange(0, W, dtype=torch.float32),
                          indexing="ij")
    xyz = torch.stack((x.reshape(-1), y.reshape(-1), torch.ones(H * W)))[None].repeat(B, 1, 1)
    rot_xyz = torch.matmul(rot, xyz)
    pxyz = rot_xyz[:, :, None, :] * depth_values.view(B, 1, D, 1) + trans.view(B, 3, 1, 1)
    pxy = pxyz[:, :2] / pxyz[:, 2:3]
    gx = pxy[:, 0] / ((W - 1) / 2) - 1
    gy = pxy[:, 1] / ((H - 1) / 2) - 1
    grid = torch.stack((gx, gy), dim=3).view(B, D * H, W, 2)
    out = F.grid_sample(src_fea, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    return out.view(B, C, D, H, W)


def _bn3d(sd, name, x):
    return F.batch_norm(x, _w(sd, f"{name}.running_mean"), _w(sd, f"{name}.running_var"),
                        _w(sd, f"{name}.weight"), _w(sd, f"{name}.bias"), training=False, eps=BN_EPS)


def cost_reg_net(sd, x, tap=_NOTAP):
    """3-D U-Net over the fused plane-sweep volume (network_v5.py:260-291; Conv3d :8-28, Deconv3d :217-258).
    Every block is conv(no bias) -> BatchNorm3d(eval) -> ReLU; skips are added after the deconv's ReLU."""
    cr = "cost_regularization"

    def conv(name, x, stride):
        y = F.conv3d(x, _w(sd, f"{cr}.{name}.conv.weight"), stride=stride, padding=1)
        return tap(f"cr.{name}", F.relu(_bn3d(sd, f"{cr}.{name}.bn", y)))

    def deconv(name, x):
        y = F.conv_transpose3d(x, _w(sd, f"{cr}.{name}.conv.weight"), stride=2, padding=1, output_padding=1)
        return F.relu(_bn3d(sd, f"{cr}.{name}.bn", y))

    c0 = conv("conv0", x, 1)
    c2 = conv("conv2", conv("conv1", c0, 2), 1)
    c4 = conv("conv4", conv("conv3", c2, 2), 1)
    x = conv("conv6", conv("conv5", c4, 2), 1)
    x = tap("cr.conv7", c4 + deconv("conv7", x))
    x = tap("cr.conv9", c2 + deconv("conv9", x))
    x = tap("cr.conv11", c0 + deconv("conv11", x))
    return tap("cr.prob", F.conv3d(x, _w(sd, f"{cr}.prob.weight"), padding=1))


def _mlp1d(sd, name, idxs, x, last_act=True):
    """Sequential of Conv1d(k=1)/Linear + ReLU at the given child indices."""
    for j, i in enumerate(idxs):
        w, b = _w(sd, f"{name}.{i}.weight"), _w(sd, f"{name}.{i}.bias")
        x = F.conv1d(x, w, b) if w.dim() == 3 else F.linear(x, w, b)
        if last_act or j + 1 < len(idxs):
            x = F.relu(x)
    return x


def ortho6d_to_mat(x_raw, y_raw):
    """rotation_utils.py:4-27: y = norm(y_raw), z = norm(x_raw x y), x = y x z, columns [x y z];
    norms clamped at 1e-8."""
    def nrm(v):
        return v / torch.clamp(torch.sqrt((v * v).sum(1, keepdim=True)), min=1e-8)
    y = nrm(y_raw)
    z = nrm(torch.cross(x_raw, y, dim=1))
    x = torch.cross(y, z, dim=1)
    return torch.stack((x, y, z), dim=2)


def decode_view(sd, feat, fused_volume, logits_volume, choose, depth_values, regress_pose=True, tap=_NOTAP):
    """Per-view decode (network_v5.py:432-465,486-499): NOCS head on the sampled pixels, softmax over the
    24 depth logits + expectation (soft-argmax), depth-guided feature fusion, pose heads."""
    B, C = feat.shape[:2]
    P = choose.shape[1]
    emb = torch.gather(feat.view(B, C, -1), 2, choose[:, None, :].repeat(1, C, 1))
    nocs_feat = _mlp1d(sd, "instance_color", (0,), emb)
    h = _mlp1d(sd, "nocs_head", (0, 2), nocs_feat)
    nocs = torch.tanh(F.conv1d(h, _w(sd, "nocs_head.4.weight"), _w(sd, "nocs_head.4.bias")))
    D = logits_volume.shape[2]
    logits = torch.gather(logits_volume.squeeze(1).reshape(B, D, -1), 2, choose[:, None, :].repeat(1, D, 1))
    prob = F.softmax(tap("logits", logits), dim=1)
    depth = torch.sum(prob * depth_values.view(B, D, 1), 1)
    out = {"nocs": nocs.permute(0, 2, 1).contiguous(), "depth": depth, "prob": prob}
    if regress_pose:
        fv = fused_volume.reshape(B, C * D, -1)
        fv = torch.gather(fv, 2, choose[:, None, :].repeat(1, C * D, 1)).view(B, C, D, P)
        fused = tap("fused_pts", torch.sum(fv * prob[:, None], dim=2))
        pts = _mlp1d(sd, "nocs_pts_mlp", (0, 2), nocs)
        pf = _mlp1d(sd, "pose_mlp1", (0, 2), torch.cat((fused, pts), dim=1))
        g = torch.mean(pf, 2, keepdim=True)
        pf2 = _mlp1d(sd, "pose_mlp2", (0, 2), torch.cat([pf, g.expand_as(pf)], 1)).mean(2)
        r6 = _mlp1d(sd, "rotation_estimator", (0, 2, 4), pf2, last_act=False)
        out["r"] = ortho6d_to_mat(r6[:, :3].contiguous(), r6[:, 3:].contiguous()).view(-1, 3, 3)
        out["t"] = _mlp1d(sd, "translation_estimator", (0, 2, 4), pf2, last_act=False)
        out["s"] = _mlp1d(sd, "size_estimator", (0, 2, 4), pf2, last_act=False)
    return out


def network_forward(sd, img1, choose1, img2, choose2, P1, P2, depth_values, regress_pose=True,
                    both_views=True, tap=_NOTAP):
    """StereoPoseNet_with_depth.forward (network_v5.py:418-519).  ``both_views=False`` skips the view-2
    cost volume / decode, which the returned box never reads under direct_regression (interface_v5.py:318-321)."""
    f1 = pspnet(sd, img1, tap)
    f2 = pspnet(sd, img2, tap)
    D = depth_values.shape[1]
    fused1 = tap("fused1", f1[:, :, None].repeat(1, 1, D, 1, 1) + homo_warping(f2, P2, P1, depth_values))
    out = {}
    v1 = decode_view(sd, f1, fused1, cost_reg_net(sd, fused1, tap), choose1, depth_values, regress_pose, tap)
    out.update({f"view1_{k}": v for k, v in v1.items()})
    if both_views:
        fused2 = f2[:, :, None].repeat(1, 1, D, 1, 1) + homo_warping(f1, P1, P2, depth_values)
        v2 = decode_view(sd, f2, fused2, cost_reg_net(sd, fused2), choose2, depth_values, regress_pose)
        out.update({f"view2_{k}": v for k, v in v2.items()})
    out["feat1"], out["feat2"] = f1, f2
    return out


# ----------------------------------------------------------------------------------------------
# transformer variant: StereoPoseNet_with_depth_baseline (ADA/lib/network_baseline.py:523-669, ADA/lib/fusion.py:11-82)
# ----------------------------------------------------------------------------------------------
def multi_head_attention(sd, name, query, key, value, heads=4):
    """fusion.py:28-51 (MultiHeadedAttention.forward; mask / dropout / position embedding are None on this path) with
    fusion.py:11-26 (scores = q k^T / sqrt(d_k), softmax over the keys).  query/key/value: [B, N, d_model]."""
    B, d = query.shape[0], query.shape[-1]
    dk = d // heads
    lin = lambda i, x: F.linear(x, _w(sd, f"{name}.linears.{i}.weight"), _w(sd, f"{name}.linears.{i}.bias"))
    q, k, v = (lin(i, x).view(B, -1, heads, dk).transpose(1, 2) for i, x in enumerate((query, key, value)))
    p = F.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk), dim=-1)
    x = torch.matmul(p, v).transpose(1, 2).contiguous().view(B, -1, d)
    return lin(3, x), p


def view_fusion(sd, f1, f2, depth=4, heads=4, tap=_NOTAP):
    """fusion.py:53-82: per block, view 1 attends to view 2 (fusion1) and view 2 to view 1 (fusion2), both from the block's
    INPUTS, each added to its own input.  f1, f2: [B, d, N]."""
    for b in range(depth):
        q, k = f1.transpose(2, 1), f2.transpose(2, 1)
        x, p1 = multi_head_attention(sd, f"view_fusion.blocks.{b}.fusion1", q, k, k, heads)
        y, _ = multi_head_attention(sd, f"view_fusion.blocks.{b}.fusion2", k, q, q, heads)
        tap(f"attn{b}", p1)
        f1, f2 = x.transpose(2, 1) + f1, y.transpose(2, 1) + f2
    return f1, f2


def network_forward_baseline(sd, img1, choose1, img2, choose2, regress_pose=True, tap=_NOTAP):
    """StereoPoseNet_with_depth_baseline.forward (network_baseline.py:606-669): PSPNet features at the sampled pixels ->
    NOCS head per view; 4 cross-view attention blocks on the raw 32-channel point features -> depth MLP (metres, ReLU) and
    the pose heads of network_v5 on cat(fused features, nocs_pts_mlp(nocs))."""
    f1, f2 = pspnet(sd, img1, tap), pspnet(sd, img2, tap)
    B, C = f1.shape[:2]
    out = {"feat1": f1, "feat2": f2}
    emb = {}
    for v, (f, ch) in enumerate(((f1, choose1), (f2, choose2)), 1):
        emb[v] = torch.gather(f.view(B, C, -1), 2, ch[:, None, :].repeat(1, C, 1)).contiguous()
        h = _mlp1d(sd, "nocs_head", (0, 2), _mlp1d(sd, "instance_color", (0,), emb[v]))
        out[f"view{v}_nocs_cf"] = torch.tanh(F.conv1d(h, _w(sd, "nocs_head.4.weight"), _w(sd, "nocs_head.4.bias")))
        out[f"view{v}_nocs"] = out[f"view{v}_nocs_cf"].permute(0, 2, 1).contiguous()
    fused = dict(zip((1, 2), view_fusion(sd, emb[1], emb[2], tap=tap)))
    for v in (1, 2):
        out[f"view{v}_fused"] = fused[v]
        out[f"view{v}_depth"] = _mlp1d(sd, "depth_head", (0, 2, 4), fused[v]).squeeze(1)
        if regress_pose:
            pts = _mlp1d(sd, "nocs_pts_mlp", (0, 2), out[f"view{v}_nocs_cf"])
            pf = _mlp1d(sd, "pose_mlp1", (0, 2), torch.cat((fused[v], pts), dim=1))
            g = torch.mean(pf, 2, keepdim=True)
            pf2 = _mlp1d(sd, "pose_mlp2", (0, 2), torch.cat([pf, g.expand_as(pf)], 1)).mean(2)
            r6 = _mlp1d(sd, "rotation_estimator", (0, 2, 4), pf2, last_act=False)
            out[f"view{v}_r"] = ortho6d_to_mat(r6[:, :3].contiguous(), r6[:, 3:].contiguous()).view(-1, 3, 3)
    return out


# ----------------------------------------------------------------------------------------------
# pose fit  (ADA/lib/utils.py:40-119, ADA/lib/align.py:10-102) and box (ADA/interface_v5.py:318-374)
# ----------------------------------------------------------------------------------------------
def back_project(depth, choose, Kp, img_size=IMG_SIZE):
    """utils.py:99-112: pixel (x = choose % S, y = choose // S) at depth z -> camera point."""
    x = (choose % img_size)[:, None]
    y = (choose // img_size)[:, None]
    z = depth[:, None]
    return np.concatenate(((x - Kp[0, 2]) * z / Kp[0, 0], (y - Kp[1, 2]) * z / Kp[1, 1], z), axis=1)


def compute_scale(cam_pts, nocs_pts):
    """utils.py:76-96: median over all ordered pairs with |dn| > 0.01 and |dc| < 0.3 of |dc| / |dn|."""
    real = np.linalg.norm(cam_pts[:, None, :] - cam_pts[None, :, :], axis=-1).flatten()
    nocs = np.linalg.norm(nocs_pts[:, None, :] - nocs_pts[None, :, :], axis=-1).flatten()
    ok = (nocs > 0.01) & (real < 0.3)
    with np.errstate(invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)   # empty selection -> NaN, like the reference
        return np.median(real[ok] / nocs[ok])


def compute_scale_and_translation(depth, nocs, choose, Kp, img_size, rotation):
    """utils.py:98-119: t = mean(cam) - mean(s R nocs)."""
    cam = back_project(depth, choose, Kp, img_size)
    s = compute_scale(cam, nocs)
    tmp = (s * rotation.astype(np.float64)) @ nocs.T.astype(np.float64)
    return cam.mean(axis=0) - tmp.T.mean(axis=0), s


def umeyama(src_h, tgt_h):
    """align.py:10-41 (similarity by SVD of the 3x3 cross-covariance)."""
    sc, tc = src_h[:3].mean(1), tgt_h[:3].mean(1)
    n = src_h.shape[1]
    cov = (tgt_h[:3] - tc[:, None]) @ (src_h[:3] - sc[:, None]).T / n
    U, Dg, Vh = np.linalg.svd(cov, full_matrices=True)
    if np.linalg.det(U) * np.linalg.det(Vh) < 0.0:
        Dg[-1] = -Dg[-1]
        U[:, -1] = -U[:, -1]
    R = U @ Vh
    scale = 1.0 / src_h[:3].var(axis=1).sum() * Dg.sum()
    t = tc - sc.dot(scale * R.T)
    T = np.identity(4)
    T[:3, :3] = scale * R
    T[:3, 3] = t
    return scale, R, t, T


def similarity_ransac(source, target, rng=np.random, rand_idx=None):
    """align.py:44-102.  ``rand_idx`` ([128,5] ints) replaces the global-RNG draws for reproducible tests."""
    n = source.shape[0]
    S = np.vstack([source.T, np.ones(n)])
    T = np.vstack([target.T, np.ones(n)])
    diam = 2 * np.amax(np.linalg.norm(S[:3] - S[:3].mean(1)[:, None], axis=0))
    inlier_t = diam / 10.0
    best_ratio, best_idx = 0, np.arange(n)
    for i in range(128):
        ridx = rng.randint(n, size=5) if rand_idx is None else rand_idx[i]
        scale, _, _, M = umeyama(S[:, ridx], T[:, ridx])
        resid = np.linalg.norm((T - M @ S)[:3], axis=0)
        idx = np.where(resid < scale * inlier_t)[0]
        ratio = idx.shape[0] / n
        if ratio > best_ratio:
            best_ratio, best_idx = ratio, idx
        if (1 - (1 - best_ratio ** 5) ** i) > 0.99:
            break
    if best_ratio < 0.1:
        return None, None, None, None
    return umeyama(S[:, best_idx], T[:, best_idx])

def prepare_pts2d(choose, rmin, rmax, cmin, img_size=IMG_SIZE):
    """interface_v5.py:136-145: image coordinates of the sampled crop pixels, x / ratio + cmin, y / ratio + rmin (float64 under
    NumPy >= 2, see prepare_model_input)."""
    ratio = img_size / (int(rmax) - int(rmin))
    x = (choose % img_size)[:, None].astype(np.float64) / ratio + int(cmin)
    y = (choose // img_size)[:, None].astype(np.float64) / ratio + int(rmin)
    return np.concatenate((x, y), axis=-1)


def triangulate_dlt(P1, P2, x1, x2):
    """cv2.triangulatePoints (OpenCV calib3d triangulate.cpp, icvTriangulatePoints): per point the 4 x 4 system
    rows x P[2] - P[0], y P[2] - P[1] of both views, solution = right singular vector of the smallest singular value.
    x1, x2: [2, N].  Returns homogeneous [4, N] (sign / norm of each column are arbitrary; callers divide by X[3])."""
    out = np.zeros((4, x1.shape[1]))
    for i in range(x1.shape[1]):
        A = np.stack([x1[0, i] * P1[2] - P1[0], x1[1, i] * P1[2] - P1[1],
                      x2[0, i] * P2[2] - P2[0], x2[1, i] * P2[2] - P2[1]])
        out[:, i] = np.linalg.svd(A)[2][3]
    return out


def nocs_matches(left_pts2d, left_nocs, left_proj, left_pose, right_pts2d, right_nocs, right_proj, right_pose, intrinsic,
                 details=None):
    """utils.py:121-195 (depth_estimation_from_nocs_matches): mutual nearest neighbours in NOCS space, distance < 0.01,
    epipolar filter < 1 px, triangulation, median scale per view.  -> (left_scale, right_scale, left_pts2d_m, right_pts2d_m)."""
    dis = np.linalg.norm(left_nocs[:, None, :] - right_nocs[None, :, :], axis=-1)
    l2r, r2l = np.argmin(dis, axis=1), np.argmin(dis, axis=0)
    lid = np.arange(left_nocs.shape[0])
    lm = lid[r2l[l2r] == lid]
    rm = l2r[lm]
    keep = dis[lm, rm] < 0.01
    lm, rm = lm[keep], rm[keep]
    rel = left_pose @ np.linalg.inv(right_pose)
    t = rel[:3, 3]
    tx = np.zeros((3, 3), np.float32)                       # float32 like the reference
    tx[0, 1], tx[1, 0], tx[0, 2], tx[2, 0], tx[1, 2], tx[2, 1] = -t[2], t[2], t[1], -t[1], -t[0], t[0]
    Ki = np.linalg.inv(intrinsic)
    f21 = Ki.T @ tx @ rel[:3, :3] @ Ki
    hl = np.vstack([left_pts2d[lm].T.astype(np.float64), np.ones(len(lm))])
    hr = np.vstack([right_pts2d[rm].T.astype(np.float64), np.ones(len(rm))])
    epi = np.abs(np.einsum("in,ij,jn->n", hl, f21, hr))     # the diagonal of hl^T f21 hr
    keep = epi < 1.0
    lm, rm = lm[keep], rm[keep]
    hl, hr = hl[:, keep], hr[:, keep]
    X = triangulate_dlt(left_proj[:3], right_proj[:3], hl[:2], hr[:2])
    X = X / X[3]
    lp, rp = left_pose @ X, right_pose @ X
    ls = compute_scale(lp[:3].T, left_nocs[lm])
    rs = compute_scale(rp[:3].T, right_nocs[rm])
    if details is not None:
        details.update(left_id=lm, right_id=rm, left_cam=lp[:3].T, right_cam=rp[:3].T, f21=f21)
    return ls, rs, left_pts2d[lm], right_pts2d[rm]


def pnp_ransac(nocs, pts2d, size, intrinsic):
    """align.py:104-115 (estimatePnPRansac): the reference's own OpenCV calls (solvePnPRansac with EPnP, 3 px, then VVS
    refinement).  The algorithm lives inside OpenCV (the image's cv2, same on the GPU box): "parity unpinned" below this call;
    what the product shares with the reference is the call itself, so the comparison is on its inputs."""
    import cv2
    tmp = nocs * size
    ok, r, t, _ = cv2.solvePnPRansac(tmp, pts2d, intrinsic, np.zeros(4), flags=cv2.SOLVEPNP_EPNP, reprojectionError=3.0)
    if ok:
        crit = (cv2.TERM_CRITERIA_MAX_ITER + cv2.TERM_CRITERIA_EPS, 20, 1e-6)
        r, t = cv2.solvePnPRefineVVS(tmp, pts2d, intrinsic, None, r, t, criteria=crit)
    R, _ = cv2.Rodrigues(r)
    return ok, size, R, t


def get_3d_bbox(size):
    """utils.py:40-58: corner order (+++,++-,-++,-+-,+-+,+--,--+,---) * size/2, returned [3,8]."""
    sx, sy, sz = size[0] / 2, size[1] / 2, size[2] / 2
    return np.array([[+sx, +sy, +sz], [+sx, +sy, -sz], [-sx, +sy, +sz], [-sx, +sy, -sz],
                     [+sx, -sy, +sz], [+sx, -sy, -sz], [-sx, -sy, +sz], [-sx, -sy, -sz]]).T


def box_from_fit(nocs, s, R, t, E1):
    """interface_v5.py:354-374: half = max|nocs|, size = 2 half s, corners -> camera (R, t; no scale in
    the matrix) -> world through inv(E1); sentinel when anything is non-finite."""
    if s is None:
        return DEFAULT_BBOX.copy()
    size = 2 * np.max(np.abs(nocs), axis=0) * s
    sRT = np.eye(4).astype(np.float32)   # the reference builds this matrix in float32
    sRT[:3, :3] = R
    sRT[:3, 3] = np.asarray(t).flatten()
    c = get_3d_bbox(size)
    cam = (sRT @ np.vstack([c, np.ones((1, 8), np.float32)]))
    cam = cam[:3] / cam[3]
    with np.errstate(all="ignore"):
        try:
            ex_inv = np.linalg.inv(E1)
        except np.linalg.LinAlgError:
            return DEFAULT_BBOX.copy()
    if np.isfinite(ex_inv).all() and np.isfinite(cam).all():
        return (ex_inv[:3, :3] @ cam + ex_inv[:3, 3:4]).T
    return DEFAULT_BBOX.copy()


# ----------------------------------------------------------------------------------------------
# end-to-end  (ADA/interface_v5.py:213-374)
# ----------------------------------------------------------------------------------------------
def predict(sd, cfg, K, rgb1, mask1, E1, rgb2, mask2, E2, rng=np.random, both_views=True, tap=_NOTAP,
            details=None, rand_idx=None):
    """One env (interface_v5.py:229-374).  ``details`` (dict) receives intermediate results."""
    S = cfg.get("img_size", IMG_SIZE)
    v1, ch1, pts1, K1 = prepare_model_input(rgb1, mask1, K, S, rng)
    v2, ch2, pts2, K2 = prepare_model_input(rgb2, mask2, K, S, rng)
    if v1 is None or v2 is None:
        return DEFAULT_BBOX.copy()
    t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()[None]
    dv = torch.from_numpy(depth_hypotheses())[None]
    regress = bool(cfg.get("direct_regression", True))
    with torch.no_grad():
        if cfg.get("name", "adapose") == "adapose_baseline":      # train.py:242-244 -> interface_baseline.py (same interface)
            pred = network_forward_baseline(sd, t32(v1), torch.from_numpy(ch1)[None], t32(v2), torch.from_numpy(ch2)[None],
                                            regress_pose=regress, tap=tap)
        else:
            pred = network_forward(sd, t32(v1), torch.from_numpy(ch1)[None], t32(v2), torch.from_numpy(ch2)[None],
                                   t32(projection(K1, E1)), t32(projection(K2, E2)), dv, regress_pose=regress,
                                   both_views=both_views, tap=tap)
    nocs = pred["view1_nocs"][0].numpy()
    depth = pred["view1_depth"][0].numpy()
    if regress:
        R = pred["view1_r"][0].numpy()
        t, s = compute_scale_and_translation(depth, nocs, ch1, K1, S, R)
    elif cfg.get("use_depth", True):
        cam = back_project(depth.flatten(), ch1, K1, S)
        s, R, t, _ = similarity_ransac(nocs, cam, rng, rand_idx)
    else:                                                   # interface_v5.py:339-349
        P1, P2 = np.eye(4), np.eye(4)
        P1[:3], P2[:3] = K @ E1[:3], K @ E2[:3]
        md = {}
        res = nocs_matches(pts1, nocs, P1, E1, pts2, pred["view2_nocs"][0].numpy(), P2, E2, K, details=md)
        _, s, R, t = pnp_ransac(nocs.astype(np.float32), pts1.astype(np.float32), res[0], K)
        if details is not None:
            details.update(matches=md, pts2d1=pts1, pts2d2=pts2, right_scale=res[1])
    if details is not None:
        details.update(view1_rgb=v1, view2_rgb=v2, choose1=ch1, choose2=ch2, K1=K1, K2=K2, nocs=nocs,
                       depth=depth, R=R, t=t, s=s, pred=pred)
    return box_from_fit(nocs, s, R, t, E1)


def estimate(sd, cfg, K, rgb1, mask1, E1, rgb2, mask2, E2, rng=np.random, both_views=True, details=None):
    """Per-env loop exactly like the reference (interface_v5.py:213-227) -> [N,8,3] float64 world boxes."""
    out = []
    for i in range(len(K)):
        d = {} if details is not None else None
        out.append(predict(sd, cfg, K[i], rgb1[i], mask1[i], E1[i], rgb2[i], mask2[i], E2[i], rng,
                           both_views=both_views, details=d))
        if details is not None:
            details.append(d)
    return np.asarray(out)


# ----------------------------------------------------------------------------------------------
# parity metrics (tolerances of BASELINE.json north_star: 0.5 px, 0.5 deg, 1 mm)
# ----------------------------------------------------------------------------------------------
def box_pose(box):
    """(centre, 3x3 axes) of an [8,3] box in the corner order of get_3d_bbox."""
    c = box.mean(axis=0)
    ax = np.stack([box[0] - box[2], box[0] - box[4], box[0] - box[1]], axis=1)  # +x, +y, +z edges
    n = np.linalg.norm(ax, axis=0)
    return c, ax / np.maximum(n, 1e-12), n


def rotation_angle_deg(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return math.degrees(math.acos(min(1.0, max(-1.0, c))))


def project_points(pts_world, K, E):
    """Pin-hole projection; also returns the camera-frame depth of every point."""
    cam = (E[:3, :3] @ pts_world.T + E[:3, 3:4])
    uv = K @ cam
    return (uv[:2] / uv[2]).T, cam[2]


def parity_errors(box_a, box_b, K, E1, min_z=0.2):
    """(max keypoint reprojection error [px], rotation error [deg], centre error [mm], max corner error [mm]).
    Keypoints = the 8 box corners and the box centre projected into view 1; points closer than ``min_z``
    metres to the camera plane (or behind it) are skipped because the projection is singular there."""
    ca, Ra, _ = box_pose(box_a)
    cb, Rb, _ = box_pose(box_b)
    pa, za = project_points(np.vstack([box_a, ca[None]]), K, E1)
    pb, zb = project_points(np.vstack([box_b, cb[None]]), K, E1)
    ok = (za > min_z) & (zb > min_z)
    px = float(np.abs(pa[ok] - pb[ok]).max()) if ok.any() else 0.0
    return (px, rotation_angle_deg(Ra, Rb), float(np.linalg.norm(ca - cb) * 1e3),
            float(np.linalg.norm(box_a - box_b, axis=1).max() * 1e3))


def keypoints_compared(box_a, box_b, K, E1, min_z=0.2):
    """(number of keypoints -- 8 corners + centre -- whose reprojection enters :func:`parity_errors`, total).  The others lie
    nearer than ``min_z`` to the camera plane in one of the two boxes and are compared in millimetres only."""
    _, za = project_points(np.vstack([box_a, box_a.mean(axis=0)[None]]), K, E1)
    _, zb = project_points(np.vstack([box_b, box_b.mean(axis=0)[None]]), K, E1)
    return int(((za > min_z) & (zb > min_z)).sum()), 9
