"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/gte_b200.h
declares, with the struct layouts the ctypes binding assumes.  No compute calls are made."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "gte_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gte_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for required in ("gte_version", "gte_last_error", "gte_reset", "gte_step", "gte_gather_obs",
                     "gte_step_obs", "gte_info"):
        assert required in names


def test_library_loads_and_exports_every_declared_symbol():
    from gym_trading_env_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        _cabi.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/gte_b200.h but not exported"
    assert set(_cabi.EXPORTS) == set(_declared_functions())
    lib.gte_version.restype = ctypes.c_int
    assert lib.gte_version() == _cabi.GTE_VERSION == 203
    lib.gte_build_id.restype = ctypes.c_char_p
    assert lib.gte_build_id().decode() == _cabi.source_hash() == _cabi.built_id()


def test_ctypes_struct_layout_matches_the_compiled_structs():
    from gym_trading_env_b200 import _cabi
    lib = _cabi.load()      # load() itself raises on an ABI mismatch
    for which, st in enumerate((_cabi.GteParams, _cabi.GteData, _cabi.GteState, _cabi.GteStepOut, _cabi.GteInfo,
                                _cabi.GteHostIO)):
        assert lib.gte_struct_size(which) == ctypes.sizeof(st), st.__name__
    assert lib.gte_struct_size(99) == -1
    assert lib.gte_default_chunks(1 << 21) >= 1


def test_bad_arguments_are_rejected_before_any_cuda_call():
    """Argument validation happens on the host: NULL structs -> GTE_ERR_ARG and a message, no launch."""
    from gym_trading_env_b200 import _cabi
    lib = _cabi.load()
    rc = lib.gte_reset(None, None, None, None, 0, None)
    assert rc == -1 and b"gte_reset" in lib.gte_last_error()
    p, d, s = _cabi.GteParams(), _cabi.GteData(), _cabi.GteState()
    rc = lib.gte_step(ctypes.byref(p), ctypes.byref(d), ctypes.byref(s), None, None, 1, None)
    assert rc == -1 and b"n_envs" in lib.gte_last_error()
    with pytest.raises(ValueError):
        _cabi.check(rc, "gte_step")


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under gym-trading-env_b200/ may import or name it."""
    pkg = os.path.join(ROOT, "gym-trading-env_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "gte_oracle" not in text and "libgte_oracle" not in text, f


def test_built_library_is_sm100a_and_uses_tma_and_mbarriers():
    """SASS evidence (B200_PROFILING.md): the gather kernel moves its windows with TMA bulk copies (UBLKCP)
    tracked by mbarriers (SYNCS), the library is built for sm_100a only, and the fp64 money math is not
    contracted into FMA in the transition kernel's portfolio arithmetic (DMUL/DADD present)."""
    import shutil
    import subprocess
    from gym_trading_env_b200 import _cabi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(_cabi.LIB_PATH):
        _cabi.build()
    elf = subprocess.run([cuobjdump, "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and "sm_90" not in elf and "sm_80" not in elf
    sass = subprocess.run([cuobjdump, "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    gather = sass[sass.index("obs_tma_coop_kernel"):]
    assert "UBLKCP" in gather                      # cp.async.bulk (TMA 1-D) global<->shared
    assert "SYNCS" in gather                       # mbarrier arrive/expect_tx/try_wait
    step = sass[sass.index("step_kernel"):]
    assert "DMUL" in step and "DADD" in step and "MUFU.RCP64H" in step   # fp64 multiply/add kept separate, IEEE divide
