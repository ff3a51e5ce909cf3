"""GPU parity for ragged batches: pairs of different sizes -- including an empty image --
go through ONE forward call with per-pair lengths (LgInputs.lens0/lens1); every pair must reproduce the
CPU oracle's answer for that pair alone (B=1, true sizes).  The reference itself cannot batch such pairs
(it pads + masks in its compiled path, lightglue.py:46-55, 256-262, 512-520)."""
import pytest
import torch

from lightglue_b200 import LightGlue, synth
from lightglue_b200.ragged import match_ragged, pad_pairs, split_outputs
from oracle import lightglue_oracle as oracle

pytestmark = pytest.mark.gpu

SIZES = [(512, 300), (200, 512), (37, 129), (0, 50), (640, 640)]


def single_pairs(sizes, *, drop_size=False, d=256, seed0=70):
    pairs = []
    for i, (m, n) in enumerate(sizes):
        data, _ = synth.make_pair(max(m, n, 1), d=d, b=1, seed=seed0 + i, m=max(m, 1))
        f0 = {k: v[:, :m].contiguous() if v.dim() == 3 else v for k, v in data["image0"].items()}
        f1 = {k: v[:, :n].contiguous() if v.dim() == 3 else v for k, v in data["image1"].items()}
        if drop_size:
            f0.pop("image_size"); f1.pop("image_size")
        pairs.append({"image0": f0, "image1": f1})
    return pairs


def cuda_pairs(pairs):
    return [{k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in p.items()} for p in pairs]


def check_pair(got, ref, *, tol, exact=True):
    flips = 0
    for k in ("matches0", "matches1"):
        a, b = got[k].cpu(), ref[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        flips += int((a != b).sum())
    if exact:
        assert flips == 0, f"{flips} index flips"
    dmax = 0.0
    for k in ("matching_scores0", "matching_scores1"):
        if ref[k].numel():
            dmax = max(dmax, float((got[k].cpu() - ref[k]).abs().max()))
    assert dmax <= tol, dmax
    assert int(got["stop"]) == int(ref["stop"]), (got["stop"], ref["stop"])
    return flips, dmax


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16x3", 1e-3)])
@pytest.mark.parametrize("drop_size", [False, True])
def test_ragged_batch_equals_per_pair_oracle(precision, tol, drop_size):
    sd = synth.make_state_dict()
    # without image_size the reference normalises by the keypoints' bounding box, undefined for an empty image
    sizes = [s for s in SIZES if not drop_size or (s[0] and s[1])]
    pairs = single_pairs(sizes, drop_size=drop_size)
    m = LightGlue(features=None, precision=precision, depth_confidence=-1, width_confidence=-1)
    m.load_state_dict(sd, strict=False)
    m = m.eval().cuda()
    results = match_ragged(m, cuda_pairs(pairs))
    assert len(results) == len(pairs)
    total = 0
    for (mm, nn), p, got in zip(sizes, pairs, results):
        ref = oracle.forward(sd, p)
        assert got["matches0"].shape == (1, mm) and got["matches1"].shape == (1, nn)
        check_pair(got, ref, tol=tol)
        if mm and nn:
            assert torch.equal(got["matches"][0].cpu().to(torch.int64), ref["matches"][0].to(torch.int64))
            total += int(ref["matches"][0].shape[0])
        else:  # the empty pair: nothing matched, no layer ran
            assert int(got["stop"]) == 1 and got["matches"][0].shape[0] == 0
            assert int((got["matches1"] != -1).sum()) == 0 and float(got["matching_scores1"].abs().max()) == 0.0
    assert total > 100  # the batch is doing real matching


def test_ragged_padding_rows_are_inert():
    """Garbage (NaN) in the padding rows must not reach any result; padding outputs are -1 / 0."""
    sd = synth.make_state_dict()
    pairs = single_pairs(SIZES[:3])
    m = LightGlue(features=None, precision="bf16x3", depth_confidence=-1, width_confidence=-1)
    m.load_state_dict(sd, strict=False)
    m = m.eval().cuda()
    data = pad_pairs(cuda_pairs(pairs))
    clean = m(data)
    lens0, lens1 = data["image0"]["num_keypoints"], data["image1"]["num_keypoints"]
    for key, lens in (("image0", lens0), ("image1", lens1)):
        for b, l in enumerate(lens.tolist()):
            data[key]["keypoints"][b, l:] = float("nan")
            data[key]["descriptors"][b, l:] = float("nan")
    dirty = m(data)
    for k in ("matches0", "matches1", "matching_scores0", "matching_scores1"):
        assert torch.equal(clean[k], dirty[k]), k
    for b, l in enumerate(lens0.tolist()):
        assert int((dirty["matches0"][b, l:] != -1).sum()) == 0 and float(dirty["matching_scores0"][b, l:].abs().sum()) == 0
    assert dirty["stops"] == clean["stops"]


def test_ragged_adaptive_matches_per_pair_oracle():
    """Early exit + point pruning on a ragged batch: per-pair stop layer, prune counters and matches
    equal the oracle's B=1 result (fp32 path; pruning from the first layer like the CPU reference)."""
    sd = synth.make_state_dict(adaptive=True)
    sizes = [(512, 384), (300, 512), (0, 64), (448, 448)]
    pairs = single_pairs(sizes, seed0=90)
    m = LightGlue(features=None, precision="fp32", depth_confidence=0.95, width_confidence=0.99)
    m.load_state_dict(sd, strict=False)
    m = m.eval().cuda()
    m.pruning_keypoint_thresholds = dict(LightGlue.pruning_keypoint_thresholds, flash=-1)
    data = pad_pairs(cuda_pairs(pairs))
    out = m(data)
    results = split_outputs(out, data["image0"]["num_keypoints"], data["image1"]["num_keypoints"])
    stops = []
    for (mm, nn), p, got in zip(sizes, pairs, results):
        ref = oracle.forward(sd, p, depth_confidence=0.95, width_confidence=0.99, pruning_threshold=-1)
        check_pair(got, ref, tol=1e-4)
        stops.append(int(ref["stop"]))
        if mm and nn:
            assert torch.equal(got["prune0"].cpu(), ref["prune0"]) and torch.equal(got["prune1"].cpu(), ref["prune1"])
    assert out["stops"] == stops
    assert out["stop"] == max(stops)
    print("ragged adaptive stops", stops)


def test_ragged_cuda_graph_replay():
    sd = synth.make_state_dict()
    m = LightGlue(features=None, precision="bf16", depth_confidence=-1, width_confidence=-1, cuda_graph=True)
    e = LightGlue(features=None, precision="bf16", depth_confidence=-1, width_confidence=-1)
    for mod in (m, e):
        mod.load_state_dict(sd, strict=False)
        mod.eval().cuda()
    for sizes in ([(256, 200), (100, 256)], [(130, 256), (256, 31)]):  # same padded shape, different lengths
        data = pad_pairs(cuda_pairs(single_pairs(sizes, seed0=120 + sizes[0][1])))
        a, b = m(data), e(data)
        for k in ("matches0", "matches1"):
            assert torch.equal(a[k], b[k])
        assert float((a["matching_scores0"] - b["matching_scores0"]).abs().max()) == 0.0
