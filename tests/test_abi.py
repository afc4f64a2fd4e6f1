"""The C-ABI library loads and exports every symbol include/cokrig.h declares; the ctypes binding
declares exactly the same set; host-only entry points behave; no compute call is made (CPU tier)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "cokrig.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ck_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = declared_symbols()
    for must in ("ck_matern_block", "ck_joint_cov", "ck_cross_cov", "ck_potrf", "ck_potrs_predict", "ck_vario_minmax",
                 "ck_vario_bin", "ck_local_predict", "ck_nll", "ck_matern_eval", "ck_distance_block"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from cokrig_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared in cokrig.h but not exported: {missing}"


def test_binding_covers_header_exactly():
    from cokrig_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_host_only_entry_points():
    from cokrig_b200 import _lib
    assert _lib.lib.ck_version() >= 100
    assert _lib.lib.ck_potrf_workspace_bytes(0) == 0
    assert _lib.lib.ck_potrf_workspace_bytes(1) == 128 * 128 * 8
    assert _lib.lib.ck_potrf_workspace_bytes(129) == 2 * 128 * 128 * 8
    assert _lib.lib.ck_vario_bin_workspace_bytes(1000, 1000, 50) > 0
    assert _lib.lib.ck_local_predict_workspace_bytes(10, 100) >= 10 * (128 + 2) * 128 * 8  # 10 slots x (kp + 2) x kp


def test_int8_path_host_only_entry_points():
    """Sizes and switches of the INT8 update path are plain host functions (no CUDA call)."""
    from cokrig_b200 import _lib
    lib = _lib.lib
    # slice buffers: 7 digit slices, A format in 128-row blocks, B format in 64-row blocks, k in 32-deep chunks
    assert lib.ck_oz_slices_bytes(128, 32, 0) == 7 * 128 * 32
    assert lib.ck_oz_slices_bytes(129, 1024, 0) == 2 * 7 * 128 * 1024
    assert lib.ck_oz_slices_bytes(64, 32, 1) == 7 * 64 * 32
    assert lib.ck_oz_slices_bytes(65, 64, 1) == 2 * 7 * 64 * 64
    assert lib.ck_oz_slices_bytes(0, 1024, 0) == 0
    assert lib.ck_oz_scales_len(1) == 128 and lib.ck_oz_scales_len(129) == 256
    # the factorisation workspace carries the slice scratch from n = 2048 on, whatever the switches say
    xinv = lambda n: ((n + 127) // 128) * 128 * 128 * 8  # noqa: E731
    assert lib.ck_potrf_workspace_bytes(2047) == xinv(2047)
    big = lib.ck_potrf_workspace_bytes(4096)
    assert big >= xinv(4096) + lib.ck_oz_slices_bytes(4096, 1024, 0) + lib.ck_oz_slices_bytes(4096, 1024, 1) + 2 * 8 * 4096
    lib.ck_oz_configure(0, -1)
    try:
        assert lib.ck_potrf_workspace_bytes(4096) == big and lib.ck_oz_active(1 << 20) == 0
    finally:
        lib.ck_oz_configure(1, -1)
    assert lib.ck_oz_active(1 << 20) == 1 and lib.ck_oz_active(2047) == 0
    # argument checks come before any CUDA call
    assert lib.ck_oz_split(None, 0, 4, 48, None, None, None, None) == _lib.CK_ERR_ARG
    assert lib.ck_oz_gemm(None, None, 4, None, None, 4, 32, None, 4, 0, 0, None) == _lib.CK_ERR_ARG


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call, with a message."""
    from cokrig_b200 import _lib
    rc = _lib.lib.ck_potrf(None, -1, 0, None, None, None)
    assert rc == _lib.CK_ERR_ARG and "negative" in _lib.last_error()
    bad = (ctypes.c_double * 11)(*([1.0] * 11))
    rc = _lib.lib.ck_joint_cov(None, 4, None, 4, bad, 3, 0, None, 8, None)
    assert rc == _lib.CK_ERR_UNSUPPORTED
    with pytest.raises(_lib.CokrigError):
        _lib.check(rc, "ck_joint_cov")


def test_no_cpu_fallback():
    """Without a CUDA device the product raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import numpy as np
    from cokrig_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.matern_eval(np.ones(3), 1.0, 1.5, 1.0)
    import model
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.MultivariateMatern().covariance(0, np.ones(3))


def test_library_contains_blackwell_native_sass():
    """The shipped library really carries the tcgen05 / TMEM / bulk-copy kernel and the FP64 DMMA kernels (SASS mnemonics:
    tcgen05.mma kind::i8 -> UTCIMMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, mma.sync f64 -> DMMA), for sm_100a only."""
    import shutil
    import subprocess
    from cokrig_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    listing = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in listing and "sm_90" not in listing and "sm_80" not in listing
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    sections, cur = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            sections[cur] = []
        elif cur is not None:
            sections[cur].append(line)
    body = lambda key: "\n".join("\n".join(v) for k, v in sections.items() if key in k)  # noqa: E731
    oz = body("ck_oz_gemm_kernel")
    for mnemonic in ("UTCIMMA", "LDTM", "UBLKCP", "UTCBAR"):
        assert mnemonic in oz, mnemonic
    assert "DMMA" in body("ck_potf2_inv_kernel") and "DMMA" in body("ck_gemm_nt_kernel")


def test_multi_gpu_handle_api_argument_checks_without_a_device():
    """ck_mg_create validates the grid before it touches CUDA or NCCL; without a device a valid request reports
    CK_ERR_CUDA (no CPU fallback); the size query of a null context is 0."""
    import torch
    from cokrig_b200 import _lib
    lib = _lib.lib
    h = ctypes.c_void_p()
    assert lib.ck_mg_create(ctypes.byref(h), 2, 0, 1, 1, 1024, None) == _lib.CK_ERR_ARG and "grid" in _lib.last_error()
    assert lib.ck_mg_create(ctypes.byref(h), 1, 0, 1, 1, 100, None) == _lib.CK_ERR_ARG and "tile" in _lib.last_error()
    assert lib.ck_mg_create(ctypes.byref(h), 4, 4, 2, 2, 1024, None) == _lib.CK_ERR_ARG
    assert lib.ck_mg_create(ctypes.byref(h), 2, 1, 1, 2, 1024, None) == _lib.CK_ERR_ARG and "unique id" in _lib.last_error()
    assert not h.value
    if not torch.cuda.is_available():
        assert lib.ck_mg_create(ctypes.byref(h), 1, 0, 1, 1, 1024, None) == _lib.CK_ERR_CUDA
    assert lib.ck_mg_workspace_bytes(None, 10, 10) == 0
    assert lib.ck_mg_destroy(None) == _lib.CK_OK
    assert lib.ck_mg_potrf(None, None) == _lib.CK_ERR_ARG


def test_header_and_c_example_compile_as_plain_c(tmp_path):
    """include/cokrig.h is a C header (C99, no C++), and examples/mg_cokrige.c -- the multi-GPU sweep driven from plain C --
    compiles and links against the library; without a GPU the program reports so and exits 1 (no CPU fallback)."""
    import shutil
    import subprocess
    from cokrig_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    src = tmp_path / "hdr.c"
    src.write_text('#include "cokrig.h"\nint main(void) { return ck_version() > 0 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", str(src)])
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA toolkit headers not found")
    exe = tmp_path / "mg_cokrige"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                           os.path.join(ROOT, "examples", "mg_cokrige.c"), "-L", libdir, "-lcokrig_b200", "-L", os.path.join(cuda, "lib64"),
                           "-lcudart", "-lm", "-o", str(exe)])
    import torch
    if not torch.cuda.is_available():
        env = dict(os.environ, LD_LIBRARY_PATH=libdir + ":" + os.path.join(cuda, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        r = subprocess.run([str(exe), "0", "1"], env=env, capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr
