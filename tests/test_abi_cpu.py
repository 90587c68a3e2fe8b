"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports every symbol the
header declares; argument validation works without a GPU (no compute is attempted here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pulse_b200 import build
    build.build()
    from pulse_b200 import _lib
    return _lib.load()


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pulse_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pulse_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    from pulse_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pulse_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in pulse_b200/_lib.py"
    assert lib.pulse_abi_version() == 1


def test_struct_sizes_match_header(lib):
    """ctypes mirrors must have the C layout: compile a tiny C program with gcc and compare sizeof."""
    import subprocess
    import tempfile
    from pulse_b200 import _lib
    src = '#include <stdio.h>\n#include "pulse_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pulse_motionlib_desc_t), sizeof(pulse_motion_query_t), sizeof(pulse_im_step_args_t), sizeof(pulse_amp_obs_args_t), sizeof(pulse_gae_args_t), sizeof(pulse_gemm_epilogue_t), sizeof(pulse_ppo_loss_args_t), sizeof(pulse_vae_latent_args_t), sizeof(pulse_reach_step_args_t));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(_lib.MotionLibDesc), C.sizeof(_lib.MotionQuery), C.sizeof(_lib.ImStepArgs), C.sizeof(_lib.AmpObsArgs),
                     C.sizeof(_lib.GaeArgs), C.sizeof(_lib.GemmEpilogue), C.sizeof(_lib.PpoLossArgs), C.sizeof(_lib.VaeLatentArgs),
                     C.sizeof(_lib.ReachStepArgs)]


def test_argument_validation_without_gpu(lib):
    from pulse_b200 import _lib
    assert lib.pulse_im_step(None, None, 4, None) == -1
    assert b"null" in lib.pulse_last_error()
    assert lib.pulse_motion_state(None, None, 1, None) == -1
    assert lib.pulse_gae(None, 32, 8, None) == -1
    a = _lib.GaeArgs()
    assert lib.pulse_gae(C.byref(a), 0, 8, None) == -1 and b"horizon" in lib.pulse_last_error()
    d = _lib.MotionLibDesc()
    h = C.c_void_p()
    assert lib.pulse_motionlib_create(C.byref(d), None, C.byref(h)) == -1
    with pytest.raises(_lib.PulseError):
        _lib.check(-1, "demo")


def test_no_cpu_fallback():
    """The product path refuses to run without a CUDA device instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pulse_b200 import PulseError
    from pulse_b200.motion_lib import MotionLibB200
    from tests.helpers import load_npz
    with pytest.raises(PulseError):
        MotionLibB200.from_tables(load_npz("motionlib.npz"))


def test_product_does_not_import_oracle():
    import glob
    for f in glob.glob(os.path.join(ROOT, "pulse_b200", "**", "*.py"), recursive=True):
        src = open(f).read()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), f"{f} references oracle/"
