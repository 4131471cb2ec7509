"""Host-side check of the warp-step decode logic (LZ4 multi-sequence + medium steps, Snappy multi-element steps).

tests/emu/*.py restate, with numpy arrays as the 32 lanes, exactly the index arithmetic the kernels in
aircompressor_b200/csrc/lz4_decode_v1.cuh and snappy.cu perform per step (candidate decode per lane, token chain, per-byte
source resolution with shuffle rounds, bail-outs to the general path).  The emulation must agree with the oracle (= Java
decoder rules) on bytes, lengths, statuses and error offsets, for valid and corrupted streams.  The GPU parity tests are the
gate for the kernels themselves; this keeps the step logic testable without a GPU."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))


def test_lz4_multisequence_step_logic(oracle):
    import lz4_multiseq_emu as emu
    checked, bad, stat = emu.main(oracle, n_cases=10)
    assert checked > 100 and bad == 0
    assert stat["seqs"] > 1.5 * stat["iters"]          # the steps really take several sequences at once
    assert stat.get("medium", 0) > 2 * stat["slow"]    # and the medium steps take most of the rest


def test_snappy_multielement_step_logic(oracle):
    import snappy_multi_emu as emu
    checked, bad, stat = emu.main(oracle, n_cases=10)
    assert checked > 100 and bad == 0
    assert stat["elems"] > 2.5 * stat["multi"]
