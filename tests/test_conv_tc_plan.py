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
    """mbarrier / commit / TMA protocol of the four csrc/conv_tc.cu kernels replayed with randomised timing
    (tools/conv_tc_protocol_sim.py): every role issues the same wait / arrive / commit sequence as the CUDA code"""
    import conv_tc_protocol_sim as S
    for issuers in (2, 1):                 # EEGB200_B1_ISSUERS: dWs UMMAs on two issuer threads (default) or one
        S.N_B1_ISSUERS = issuers
        for mode in ("B1", "B2"):
            for gc in (4, 3):
                for n_tiles in (1, 2, 3, 7):
                    for seed in range(6):
                        ok, stuck = S.run(mode, gc, n_tiles, seed)
                        assert ok, (issuers, mode, gc, n_tiles, seed, stuck)
    S.N_B1_ISSUERS = 2
    for n_items in (1, 3):
        for seed in range(8):
            ok, stuck = S.run_fwd(n_items, seed)
            assert ok, (n_items, seed, stuck)


def test_attention_block_diagonal_plan_matches_autograd():
    """two samples of a head as one 128-row UMMA problem (attention_tc.cu forward; backward = round-2 plan)"""
    import attn_tc_plan
    assert attn_tc_plan.main(seed=0) == 0 and attn_tc_plan.main(seed=3, p_drop=0.0) == 0
