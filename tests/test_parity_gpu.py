"""GPU parity tests proper: every check calls libecamp_b200.so through the C ABI (ctypes) or through the
drop-in nn.Module and compares with the oracle / golden fixtures.  See tests/parity_checks.py for the
checks and the tolerances."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pc():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    from tests import parity_checks
    return parity_checks


@pytest.mark.parametrize("name", ["check_masking", "check_resize", "check_gemm", "check_layernorm", "check_attention", "check_losses",
                                  "check_ce", "check_step", "check_adamw", "check_finetune_cls", "check_sgd", "check_attention_map", "check_image_u8", "check_full_size_batch_split", "check_edge_cases", "check_stage_final", "check_trainer_recipe",
                                  "check_optimizer_resume", "check_step_full_size", "check_gemm_fp32", "check_step_fp32", "check_image_pipeline", "check_adamw_groups", "check_side_stream", "check_overlapped_update", "check_dropout_gradient"])
def test_parity(pc, name, capsys):
    ok = pc.run_check(getattr(pc, name))
    out = capsys.readouterr().out
    assert ok, out[-4000:]


def test_native_library_is_what_runs(pc):
    """The product path must be the CUDA library: kernels were launched, and a missing library fails loudly."""
    from ecamp_b200 import _lib
    assert _lib.lib().ecamp_launch_count() > 0
    import ecamp_b200._lib as L
    saved, L._lib, path = L._lib, None, L.LIB_PATH
    L.LIB_PATH = path + ".missing"
    try:
        with pytest.raises(RuntimeError):
            L.lib()
    finally:
        L._lib, L.LIB_PATH = saved, path
