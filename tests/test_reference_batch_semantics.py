"""What the reference does with adaptive depth on a batch of more than one pair -- run on the UNMODIFIED reference file
(oracle/_ref/lightglue_ref.py, CPU) -- and therefore why lightglue_b200 decides per pair (documented deviation,
LightGlue.forward docstring).

lightglue.py:645-656 (`check_if_stop`): the low-confidence count is summed over the WHOLE batch and divided by ONE pair's
m + n, and one decision is taken for all pairs.  So a pair's result depends on its batch-mates -- even on being duplicated:
two copies of a pair that stops early alone run all nine layers together, with different matches.  (With point pruning on,
`torch.where(mask)[1]` at 554 / 562 additionally concatenates the kept columns of all rows: not defined for B > 1.)"""
import pytest
import torch

from lightglue_b200 import synth
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref/lightglue_ref.py not built (make -C oracle ref)")

WEIGHT_SEED, N = 2, 192  # weights with which pair 41 exits early and pair 42 does not


def _cat(pairs):
    return {k: {kk: torch.cat([p[k][kk] for p in pairs]) for kk in pairs[0][k]} for k in ("image0", "image1")}


def test_reference_batched_early_exit_depends_on_batch_mates():
    torch.set_grad_enabled(False)
    sd = synth.make_state_dict(adaptive=True, seed=WEIGHT_SEED)
    ref = ref_loader.build_matcher(sd, depth_confidence=0.95, width_confidence=-1)
    p41, p42 = (synth.make_pair(N, b=1, seed=s)[0] for s in (41, 42))
    alone41, alone42 = ref(p41), ref(p42)
    assert int(alone41["stop"]) < int(alone42["stop"]) == 9  # one pair exits early, the other never
    twice = ref(_cat([p41, p41]))                            # the SAME pair, twice in one batch
    assert int(twice["stop"]) == 9 > int(alone41["stop"])    # ... no longer exits: count summed over the batch / one pair's m + n
    assert not torch.equal(twice["matches0"][0], alone41["matches0"][0])  # and its matches changed with it
    assert torch.equal(twice["matches0"][0], twice["matches0"][1])
    mixed = ref(_cat([p41, p42]))
    assert int(mixed["stop"]) == 9  # one batch-global decision: pair 41 is dragged along
