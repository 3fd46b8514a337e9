"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/consenrich_b200.h declares, the ctypes table matches the header, the
host-side mirror validates arguments with the reference's messages before any device work, and
nothing on the product path can run without a GPU (no CPU fallback, no oracle import)."""
import ast
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "consenrich_b200.h")


def _declared():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"CB200_API[^;(]*?\b(cb200_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    from consenrich_b200 import build
    path = build.build()
    assert os.path.exists(path)
    return path


def test_header_declares_the_six_reference_entry_points():
    names = _declared()
    assert len(names) >= 30
    for need in ("cb200_host_forward_pass", "cb200_host_backward_pass", "cb200_host_ecm", "cb200_fold_tracks",
                 "cb200_forward_scan", "cb200_backward_scan", "cb200_residuals"):
        assert need in names
    text = open(HEADER).read()
    for cite in ("6393", "6635", "6853", "7052", "7153", "7660", "259-283"):
        assert cite in text  # every entry point cites the reference lines it replaces


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", built_lib], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in _declared() if n not in exported]
    assert not missing, missing
    # nothing but the C ABI leaks out of the library
    assert all(n.startswith("cb200_") for n in exported), sorted(n for n in exported if not n.startswith("cb200_"))


def test_library_carries_sm100a_code_only(built_lib):
    out = subprocess.check_output(["cuobjdump", "--list-elf", built_lib], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_ctypes_table_matches_header(built_lib):
    from consenrich_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.cb200_abi_version() == _lib.ABI_VERSION
    # struct sizes agree with the C definitions (compile a probe with the real header)
    probe = os.path.join(ROOT, "tests", "emul", "_build", "abi_probe")
    os.makedirs(os.path.dirname(probe), exist_ok=True)
    src = probe + ".c"
    with open(src, "w") as f:
        f.write('#include <stdio.h>\n#include "consenrich_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                "sizeof(cb200_model),sizeof(cb200_ecm_opts),sizeof(cb200_ecm_result),"
                "sizeof(cb200_munc_finalize_result),sizeof(cb200_munc_seed_args));return 0;}\n")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", probe])
    sizes = [int(x) for x in subprocess.check_output([probe], text=True).split()]
    import ctypes as C
    assert sizes == [C.sizeof(_lib.Model), C.sizeof(_lib.EcmOpts), C.sizeof(_lib.EcmResult),
                     C.sizeof(_lib.MuncFinalizeResult), C.sizeof(_lib.MuncSeedArgs)]


def test_no_device_means_loud_failure_not_a_cpu_path(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from consenrich_b200 import _lib
    with pytest.raises(_lib.CudaError, match="no CPU path"):
        _lib.Context(0)
    import consenrich_b200 as cb
    data = np.zeros((2, 16), np.float32)
    munc = np.ones((2, 16), np.float32)
    with pytest.raises(_lib.CudaError):
        cb.cforwardPass(matrixData=data, matrixPluginMuncInit=munc, matrixF=np.eye(2, dtype=np.float32),
                        matrixQ0=np.eye(2, dtype=np.float32) * 1e-3, intervalToBlockMap=np.zeros(16, np.int32),
                        blockCount=1, stateInit=0.0, stateCovarInit=1000.0)


def test_argument_validation_precedes_device_work(built_lib):
    """ValueErrors carry the reference's texts (cconsenrich.pyx:6503-6561) and fire without a GPU."""
    import consenrich_b200 as cb
    data = np.zeros((2, 16), np.float32)
    munc = np.ones((2, 16), np.float32)
    base = dict(matrixData=data, matrixPluginMuncInit=munc, matrixF=np.eye(2, dtype=np.float32),
                matrixQ0=np.eye(2, dtype=np.float32) * 1e-3, intervalToBlockMap=np.zeros(16, np.int32),
                blockCount=1, stateInit=0.0, stateCovarInit=1000.0)
    with pytest.raises(ValueError, match="blockCount must be positive"):
        cb.cforwardPass(**{**base, "blockCount": 0})
    with pytest.raises(ValueError, match="matrixPluginMuncInit shape must match matrixData shape"):
        cb.cforwardPass(**{**base, "matrixPluginMuncInit": munc[:, :8].copy()})
    with pytest.raises(ValueError, match=r"processQScale\[0\] must be 1.0"):
        cb.cforwardPass(**base, processQScale=np.full(16, 2.0, np.float32))
    with pytest.raises(ValueError, match="process precision multiplier bounds"):
        cb.cforwardPass(**base, procPrecisionMultiplierMax=0.1)
    with pytest.raises(ValueError, match=r"matrixQ0\[0, 0\] must be positive"):
        cb.cforwardPassLevel(**{k: v for k, v in base.items() if k != "matrixF"} | {"matrixQ0": np.zeros((1, 1), np.float32)})
    with pytest.raises(ValueError, match="matrixQ0 is singular"):
        cb.cfixedBackgroundECM(**{**base, "matrixQ0": np.ones((2, 2), np.float32)})
    # adaptive process noise is a device path like any other (csrc/apn_kernels.cu): without a device it fails
    # loudly, it does not fall back
    import torch
    from consenrich_b200 import _lib
    if not torch.cuda.is_available():
        with pytest.raises((_lib.CudaError, _lib.NativeLibraryMissing)):
            cb.cforwardPass(**base, ECM_useAPN=True)
    # empty input: zeros, no device needed (pyx:6494-6501)
    e = np.empty((2, 0), np.float32)
    r = cb.cforwardPass(**{**base, "matrixData": e, "matrixPluginMuncInit": e, "intervalToBlockMap": np.zeros(0, np.int32)})
    assert r[0] == 0.0 and r[1] == 0 and r[2].shape == (0,)


def test_signatures_mirror_the_reference_keywords():
    """Keyword names and defaults of the six functions (cconsenrich.pyx:6393-6428, 6635-6646,
    6853-6884, 7052-7062, 7153-7185, 7660-7693), checked against the oracle's restatement."""
    import inspect

    import consenrich_b200.native as N
    from oracle import oracle as O
    for name in N._HOT_PATH:
        a, b = inspect.signature(getattr(N, name)), inspect.signature(getattr(O, name))
        assert list(a.parameters) == list(b.parameters), name
        for p in a.parameters:
            assert a.parameters[p].default == b.parameters[p].default, (name, p)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under consenrich_b200/ may import or open it."""
    pkg = os.path.join(ROOT, "consenrich_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            path = os.path.join(dirpath, fn)
            if fn.endswith(".py"):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    mods = []
                    if isinstance(node, ast.Import):
                        mods = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        mods = [node.module or ""]
                    assert not any(m.split(".")[0] == "oracle" for m in mods), path
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                assert "ssm_oracle" not in open(path).read(), path
