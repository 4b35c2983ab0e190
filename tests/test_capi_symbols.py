"""CPU: the C-ABI shared library builds/loads here (no GPU) and exports every
symbol include/histogan_b200.h declares; host-only entry points work."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            src = open(os.path.join(inc, f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from histogan_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert set(_lib.exported_symbols()) == set(declared), \
        "ctypes binding table and header disagree"
    assert lib.hg_abi_version() == _lib.HG_ABI_VERSION


def test_host_only_geometry():
    from histogan_b200 import _lib
    lib = _lib.load()

    def npix(H, W, insz, resizing, h=64):
        p = _lib.HistParams(2, 3, H, W, 3 * H * W, H * W, W, 1, h, insz, resizing, 2, 0.02,
                            -3.0, 3.0, 1, 0)
        return lib.hg_hist_num_pixels(C.byref(p))

    assert npix(64, 64, 150, 0) == 64 * 64            # no resize      (RGBuvHistBlock.py:77)
    assert npix(256, 256, 150, 0) == 150 * 150        # interpolation  (:78-80)
    assert npix(256, 256, 150, 1) == 64 * 64          # sampling -> h x h (:81-89)
    assert npix(100, 200, 150, 0) == 150 * 150        # one side larger is enough
    bad = _lib.HistParams(2, 2, 8, 8, 128, 64, 8, 1, 64, 150, 0, 2, 0.02, -3.0, 3.0, 1, 0)
    assert lib.hg_hist_num_pixels(C.byref(bad)) < 0   # C < 3
    assert b"C>=3" in lib.hg_last_error()


def test_block_rejects_cpu_and_bad_args():
    import torch
    from histogan_b200 import RGBuvHistBlock
    blk = RGBuvHistBlock(device="cpu")
    assert list(blk.state_dict().keys()) == []        # no parameters / buffers
    with pytest.raises(RuntimeError, match="CUDA"):
        blk(torch.rand(1, 3, 8, 8))
    b = [3, -3]
    RGBuvHistBlock(hist_boundary=b)
    assert b == [-3, 3]                               # sorted in place (RGBuvHistBlock.py:68)


def test_struct_layouts_match_the_header(tmp_path):
    """the ctypes mirrors of the ABI structs have the size / field offsets a C compiler gives the
    declarations in include/histogan_b200.h (compiled here with gcc as plain C)."""
    import shutil
    import subprocess
    from histogan_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "histogan_b200.h"
int main(void) {
  printf("%zu %zu %zu\n", sizeof(hg_hist_params), offsetof(hg_hist_params, sigma), offsetof(hg_hist_params, projection));
  printf("%zu %zu\n", sizeof(hg_conv_params), offsetof(hg_conv_params, OW));
  printf("%zu %zu %zu %zu\n", sizeof(hg_conv_epilogue), offsetof(hg_conv_epilogue, noise_size),
         offsetof(hg_conv_epilogue, lrelu_slope), offsetof(hg_conv_epilogue, out_pix_stride));
  printf("%d\n", HG_ABI_VERSION);
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = list(map(int, out))
    H, P, E = _lib.HistParams, _lib.ConvParams, _lib.ConvEpilogue
    want = [C.sizeof(H), H.sigma.offset, H.projection.offset,
            C.sizeof(P), P.OW.offset,
            C.sizeof(E), E.noise_size.offset, E.lrelu_slope.offset, E.out_pix_stride.offset,
            _lib.HG_ABI_VERSION]
    assert got == want, (got, want)
