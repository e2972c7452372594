"""The round-2 blocking of the conv stack (docs/ROUND2_CONV_TCGEN05.md, tools/conv_tc_plan.py) reproduces autograd through
the reference's torch ops (ATMS_retrieval.py:101-106): pooled-sum im2col tiles, 3xTF32 conv product, per-channel
streaming of the spatial contraction, folded BatchNorm-backward constants, prefix-scan transposed pooling."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_conv_tc_plan_matches_autograd():
    import conv_tc_plan
    assert conv_tc_plan.main(B=4, seed=0) == 0


def test_conv_tc_plan_ragged_last_tile():
    import conv_tc_plan
    assert conv_tc_plan.main(B=5, seed=1) == 0        # 5 = 3 + 2: the second tile has 72 valid rows


def test_conv_tc_kernel_protocol_has_no_deadlock_or_hazard():
    """mbarrier / commit / TMA protocol of csrc/conv_tc.cu replayed with randomised timing (tools/conv_tc_protocol_sim.py)"""
    import random
    import conv_tc_protocol_sim as S
    for mode in ("stats", "apply"):
        for seed in range(40):
            S.Sim(mode, random.Random(seed)).run()


def test_attention_block_diagonal_plan_matches_autograd():
    """two samples of a head as one 128-row UMMA problem (attention_tc.cu forward; backward = round-2 plan)"""
    import attn_tc_plan
    assert attn_tc_plan.main(seed=0) == 0 and attn_tc_plan.main(seed=3, p_drop=0.0) == 0
