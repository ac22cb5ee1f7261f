"""The multi-threaded torch CPU restatement used as the reported CPU baseline (oracle/cpu_path.py) agrees
with the plain-C oracle (which is pinned to outputs of the reference). No GPU."""
import numpy as np
import torch

from oracle import cpu_path as P
from oracle import oracle as O


def test_pool_bit_equal(golden):
    g = golden("pool_a")
    out = P.voxel_maxpool(torch.from_numpy(g["feat"]), torch.from_numpy(g["ind"]), (int(g["H"]), int(g["W"])),
                          tuple(float(s) for s in g["scale"]))
    assert np.array_equal(out.numpy(), g["out"])


def test_bilinear_matches_golden(golden):
    g = golden("bilinear_a")
    out = P.bilinear_sample(torch.from_numpy(g["grid"]), torch.from_numpy(g["coord"]), tuple(float(s) for s in g["scale"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-6, atol=1e-7)


def test_msda_matches_golden(golden):
    g = golden("msda_streammos_small")
    out = P.ms_deform_attn(torch.from_numpy(g["value"]).double(), [tuple(r) for r in g["shapes"]],
                           torch.from_numpy(g["loc"]).double(), torch.from_numpy(g["attn"]).double())
    np.testing.assert_allclose(out.numpy(), g["out64"], rtol=1e-10, atol=1e-12)


def test_voting_matches_golden(golden):
    g = golden("voting_a")
    size = tuple(int(s) for s in g["size"])
    q = P.quantize(torch.from_numpy(g["pts"]), tuple(g["rx"]), tuple(g["ry"]), tuple(g["rz"]), size)
    assert np.array_equal(q.numpy(), g["quan"])
    vl = P.determine_voxel_labels(q.to(torch.int64), torch.from_numpy(g["labels"]), size)
    assert np.array_equal(vl.numpy(), g["voxel_labels"])
    pl = P.get_point_labels_from_voxel_labels(torch.from_numpy(g["cur"]), vl, size)
    assert np.array_equal(pl.numpy(), g["point_labels"])


def test_instance_vote_matches_golden(golden):
    g = golden("instance_a")
    lo, hi = torch.from_numpy(g["corners"].min(1)), torch.from_numpy(g["corners"].max(1))
    sums = P.instance_vote(torch.from_numpy(g["local_pts"]), torch.from_numpy(g["local_pred"]), lo, hi)
    assert np.array_equal(sums[:, 0], g["stat"]) and np.array_equal(sums[:, 1], g["dyn"])
    assert np.array_equal(sums, O.instance_vote(g["local_pts"], g["local_pred"], lo.numpy(), hi.numpy()))
