"""Multi-threaded CPU restatement of the hot path with torch CPU ops — the REPORTED CPU BASELINE.

TEST/BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py). This is the path BASELINE.json names as the
reference's CPU path: "the repo's ms_deform_attn_core_pytorch, numpy voting and a torch scatter_reduce
equivalent of the deep_point pooling" (BASELINE.md §3). /root/reference cannot travel to the GPU box, so
the algorithms are restated here; tests/test_cpu_path.py checks each against oracle/smos_oracle.c, which
is itself pinned to outputs of the reference (tests/golden).

  pool      : scatter_reduce_('amax', include_self=False) over a zero grid with a dump slot — bit-equal to
              deep_point/src/point_deep.cpp:19-88
  gather    : networks/backbone.py:458-475 (normalise + F.grid_sample bilinear/zeros/align_corners=True)
  msda      : deformattn/functions/ms_deform_attn_func.py:41-61
  voting    : voxel_voting.py:38-91 (torch, CPU tensors)
  instance  : voxel_instance_voting.py:177-187 with in_hull restated as the inclusive AABB test on torch CPU
              ops (far faster than the reference's per-cluster scipy Delaunay, i.e. the baseline is favoured)
"""
import numpy as np
import torch
import torch.nn.functional as F


def voxel_maxpool(feat, ind, output_size, scale_rate):
    """feat (B,C,N,1), ind (B,N,2,1) -> (B,C,H,W)."""
    B, C, N = feat.shape[:3]
    H, W = output_size
    f = feat.reshape(B, C, N)
    ih = (ind[:, :, 0, 0] * scale_rate[0]).to(torch.int64)  # fp32 product, truncation toward zero
    iw = (ind[:, :, 1, 0] * scale_rate[1]).to(torch.int64)
    valid = (ih >= 0) & (ih < H) & (iw >= 0) & (iw < W)
    cell = torch.where(valid, ih * W + iw, torch.full_like(ih, H * W))  # invalid -> dump slot
    out = torch.zeros(B, C, H * W + 1, dtype=feat.dtype)
    out.scatter_reduce_(2, cell[:, None, :].expand(B, C, N), f, reduce="amax", include_self=False)
    return out[:, :, : H * W].reshape(B, C, H, W)


def bilinear_sample(grid_feat, grid_coord, scale_rate):
    """grid_feat (B,C,H,W), grid_coord (B,N,2,S) -> (B,C,N,S)."""
    H, W = grid_feat.shape[2:]
    gx = 2 * grid_coord[:, :, 1] * scale_rate[1] / (W - 1) - 1
    gy = 2 * grid_coord[:, :, 0] * scale_rate[0] / (H - 1) - 1
    return F.grid_sample(grid_feat, torch.stack((gx, gy), dim=-1), mode="bilinear", padding_mode="zeros",
                         align_corners=True)


def ms_deform_attn(value, spatial_shapes, sampling_locations, attention_weights):
    """value (B,S,M,D), loc (B,Q,M,L,P,2), attn (B,Q,M,L,P) -> (B,Q,M*D)."""
    B, S, M, D = value.shape
    _, Q, _, L, P, _ = sampling_locations.shape
    sizes = [int(h) * int(w) for h, w in spatial_shapes]
    grids = 2 * sampling_locations - 1
    sampled = []
    for lvl, v in enumerate(value.split(sizes, dim=1)):
        h, w = (int(x) for x in spatial_shapes[lvl])
        v = v.flatten(2).transpose(1, 2).reshape(B * M, D, h, w)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)  # (B*M, Q, P, 2)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    a = attention_weights.transpose(1, 2).reshape(B * M, 1, Q, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * a).sum(-1).view(B, M * D, Q)
    return out.transpose(1, 2).contiguous()


def quantize(pcds, range_x, range_y, range_z, size):
    d = [(r[1] - r[0]) / s for r, s in zip((range_x, range_y, range_z), size)]
    return torch.stack(((pcds[:, 0] - range_x[0]) / d[0], (pcds[:, 1] - range_y[0]) / d[1],
                        (pcds[:, 2] - range_z[0]) / d[2]), dim=-1)


def determine_voxel_labels(voxel_coords, semantic_labels, size):
    num_classes = int(semantic_labels.max().item()) + 1
    X, Y, Z = size
    inside = ((voxel_coords >= 0).all(1) & (voxel_coords[:, 0] < X) & (voxel_coords[:, 1] < Y) &
              (voxel_coords[:, 2] < Z))  # the reference crops before (transforms.py:151-161); pads are dropped here
    vc, sl = voxel_coords[inside], semantic_labels[inside]
    lin = vc[:, 0] * Y * Z + vc[:, 1] * Z + vc[:, 2]
    votes = torch.zeros(X * Y * Z, num_classes, dtype=torch.long)
    votes.scatter_add_(0, lin[:, None].expand(-1, num_classes), F.one_hot(sl, num_classes))
    return votes.view(X, Y, Z, num_classes).argmax(dim=-1)


def get_point_labels_from_voxel_labels(coords, voxel_labels, size):
    X, Y, Z = size
    ok = (coords >= 0).all(1) & (coords[:, 0] < X) & (coords[:, 1] < Y) & (coords[:, 2] < Z)
    out = torch.zeros(coords.shape[0], dtype=torch.long)
    c = coords[ok]
    out[ok] = voxel_labels.reshape(-1)[c[:, 0] * Y * Z + c[:, 1] * Z + c[:, 2]]
    return out


def instance_vote(points, pred, box_lo, box_hi):
    """torch (multi-threaded) inclusive AABB test per box; sums[k] = [count(pred==1), 2*count(pred==2)]."""
    xyz = points[:, :3]
    is1, is2 = pred == 1, pred == 2
    sums = torch.zeros((box_lo.shape[0], 2), dtype=torch.int64)
    for k in range(box_lo.shape[0]):
        inside = ((xyz >= box_lo[k]) & (xyz <= box_hi[k])).all(dim=1)
        sums[k, 0] = (inside & is1).sum()
        sums[k, 1] = 2 * (inside & is2).sum()
    return sums.numpy()


class CpuHotPath:
    """Same per-scan sequence as streammos_b200.stream.HotPath.step, on CPU tensors."""

    def __init__(self, hot):
        """`hot`: a dict of CPU tensors with the resident state (x0, x1, dec, memory, local_pts, local_pred,
        box_lo, box_hi) — see bench.py."""
        self.s = hot

    def step(self, b):
        s = self.s
        cur_bev, cur_rv = b.coord_bev[:1], b.coord_rv
        bev_in = voxel_maxpool(b.feat, b.coord_bev, (512, 512), (1.0, 1.0))
        x0_pt = bilinear_sample(s["x0"], cur_bev, (0.5, 0.5))
        x0_rv = voxel_maxpool(x0_pt, cur_rv, (32, 1024), (0.5, 0.5))
        x0_pt = bilinear_sample(x0_rv, cur_rv, (0.5, 0.5))
        x0_bev = voxel_maxpool(x0_pt, cur_bev, (256, 256), (0.5, 0.5))
        x1_pt = bilinear_sample(s["x1"], cur_bev, (0.25, 0.25))
        x1_rv = voxel_maxpool(x1_pt, cur_rv, (16, 512), (0.25, 0.25))
        x1_pt = bilinear_sample(x1_rv, cur_rv, (0.25, 0.25))
        x1_bev = voxel_maxpool(x1_pt, cur_bev, (128, 128), (0.25, 0.25))
        pt_bev = bilinear_sample(s["dec"], cur_bev, (0.5, 0.5))
        value = s["memory"].view(1, 4096, 4, 32)
        h = ms_deform_attn(value, [(64, 64)], b.loc[0], b.attn[0])
        h = ms_deform_attn(h.view(1, 4096, 4, 32), [(64, 64)], b.loc[1], b.attn[1])
        s["memory"] = h
        n = b.xyzi.shape[0]
        s["local_pts"][8].copy_(b.xyzi)
        s["local_pred"][8].copy_(b.pred)
        pts = s["local_pts"].view(-1, 4)
        size = (512, 512, 30)
        q = quantize(pts, (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), size)
        coords = q.to(torch.int64)
        labels = s["local_pred"].view(-1).to(torch.int64)
        vl = determine_voxel_labels(coords, labels, size)
        pl = get_point_labels_from_voxel_labels(coords[8 * n:], vl, size)
        sums = instance_vote(pts, labels, s["box_lo"], s["box_hi"])
        slot = s["scan_index"] % 8
        s["local_pts"][slot].copy_(s["local_pts"][8])
        s["local_pred"][slot].copy_(s["local_pred"][8])
        s["scan_index"] += 1
        return pl, sums, (bev_in, x0_bev, x1_bev, x1_pt, pt_bev)
