"""CPU: the C-ABI library builds, loads, and exports every symbol that
include/pcadv.h declares (no compute calls -- there is no GPU here); the ctypes
mirrors have the C struct sizes; the product path fails loudly without CUDA."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest
import torch

from adversarial_learning_on_pointclouds_b200 import _lib, _build, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcadv.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcadv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.pcadv_version() == 100
    assert lib.pcadv_launch_count() >= 0


def test_ctypes_structs_match_c_layout():
    code = '#include <stdio.h>\n#include "pcadv.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",' \
           'sizeof(pcadv_seg),sizeof(pcadv_linear_args),sizeof(pcadv_wgrad_args),' \
           'sizeof(pcadv_maxbwd_args),sizeof(pcadv_backlevel_args),sizeof(pcadv_chain_args),' \
           'sizeof(pcadv_head_args));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(code)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.Seg), ctypes.sizeof(_lib.LinearArgs),
                     ctypes.sizeof(_lib.WgradArgs), ctypes.sizeof(_lib.MaxBwdArgs),
                     ctypes.sizeof(_lib.BackLevelArgs), ctypes.sizeof(_lib.ChainArgs),
                     ctypes.sizeof(_lib.HeadArgs)]


def test_header_is_plain_c():
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "h.c")
        open(c, "w").write('#include "pcadv.h"\nint main(void){return PCADV_VERSION == 100 ? 0 : 1;}\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c,
                        "-o", os.path.join(td, "h")], check=True)


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    a = _lib.LinearArgs()
    assert lib.pcadv_linear(ctypes.byref(a), None) != 0
    assert b"pcadv_linear" in lib.pcadv_last_error()
    w = _lib.WgradArgs()
    assert lib.pcadv_wgrad(ctypes.byref(w), None) != 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    with pytest.raises(RuntimeError):
        ops.linear([torch.zeros(4, 4)], torch.zeros(4, 4))
    with pytest.raises(_lib.PcadvError):
        _lib.lib()


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", _build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
