"""CPU tests of the drop-in boundary: the shared library loads without a GPU, exports every symbol
include/*.h declares, keeps the reference's struct layout, and fails loudly (NULL + message, no CPU
fallback) when no CUDA device is usable."""
import ctypes as C
import re
from pathlib import Path

import pytest

import __graft_entry__ as entry

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    pkg = entry.load_package()
    lib = pkg.load()
    declared = set()
    for header in (ROOT / "include").glob("*.h"):
        text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
        declared |= set(re.findall(r"\b((?:resample|biquad_|decimate|floatIntegers)\w+)\s*\(", text))
    assert declared, "no prototypes found in include/"
    assert declared == set(pkg.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"


def test_struct_layout_matches_reference():
    pkg = entry.load_package()
    r = pkg.Resample
    # offsets measured on the reference build
    assert (r.numChannels.offset, r.numSamples.offset, r.numFilters.offset, r.numTaps.offset,
            r.inputIndex.offset, r.flags.offset) == (0, 4, 8, 12, 16, 20)
    assert (r.tempFilter.offset, r.outputOffset.offset, r.fixedRatio.offset, r.lowpassRatio.offset,
            r.subsample.offset, r.buffers.offset, r.filters.offset) == (24, 32, 40, 48, 56, 64, 72)
    assert C.sizeof(pkg.Biquad) == 80 and C.sizeof(pkg.ResampleResult) == 8


def test_init_argument_validation_matches_reference(capfd):
    lib = entry.load_package().load()
    assert not lib.resampleInit(2, 30, 48, 0.0, 3)          # taps not a multiple of 4 (resampler.c:127)
    assert not lib.resampleInit(2, 2048, 48, 0.0, 3)
    assert not lib.resampleInit(2, 48, 0, 0.0, 3)           # filters out of range (resampler.c:132)
    assert not lib.resampleInit(2, 48, 2000, 0.0, 3)
    assert not lib.resampleFixedRatioInit(2, 48, 48, 44100.0, 48000.0, 30000, 7)   # lowpass above Nyquist (:316)
    err = capfd.readouterr().err
    assert "multiple of 4" in err and "1-1024 filters" in err and "destination Nyquist" in err
    lib.resampleFree(None)                                   # accepts NULL (resampler.c:975)


def test_no_gpu_means_null_not_fallback(capfd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = entry.load_package().load()
    assert not lib.resampleInit(2, 48, 48, 0.0, 3)
    assert "no usable CUDA device" in capfd.readouterr().err
    assert lib.resampleB200GetDeviceCount() == 0


def test_biquad_design_runs_on_host():
    import numpy as np
    pkg = entry.load_package()
    lib = pkg.load()
    co = pkg.BiquadCoefficients()
    lib.biquad_lowpass(C.byref(co), 0.45 * 44100 / 96000)
    got = np.array([co.a0, co.a1, co.a2, co.b1, co.b2], np.float32)
    want = np.load(ROOT / "tests" / "golden" / "biquad_lowpass_0p2067.npz")["coeffs"]
    assert np.array_equal(got, want)
    q = pkg.Biquad()
    lib.biquad_init(C.byref(q), C.byref(co), 1.0)
    assert q.order == 2 and q.index == 0


@pytest.mark.parametrize("width", [32, 64])
def test_c_caller_compiles_and_links_against_the_headers(tmp_path, width):
    """A plain C caller, compiled against include/*.h exactly as a caller of the reference would be (-DPATH_WIDTH=64 selects the
    double-sample build, reference resampler.h:22-26), links against the matching library and sees the reference's struct layout."""
    import subprocess
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdio.h>
#include "resampler.h"
#include "biquad.h"
#include "decimator.h"
#include "resampler_b200.h"
int main (void)
{
    Resample *r;
    BiquadCoefficients c;
    printf ("%d %d %d %d %d\n", (int) sizeof (artsample_t), (int) offsetof (Resample, outputOffset), (int) offsetof (Resample, filters),
            (int) sizeof (Biquad), (int) sizeof (ResampleResult));
    biquad_lowpass (&c, 0.2);                       /* host-side design: works without a GPU */
    printf ("%d\n", c.a0 > 0 && c.a1 > c.a0);
    r = resampleInit (2, 48, 48, 0.0, SUBSAMPLE_INTERPOLATE | BLACKMAN_HARRIS);      /* NULL + message when no CUDA device is usable */
    printf ("%d\n", r != NULL);
    resampleFree (r);
    return 0;
}
''')
    lib = "resampler_b200_64" if width == 64 else "resampler_b200"
    libdir = ROOT / "audio-resampler_b200" / "lib"
    exe = tmp_path / "caller"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(src), f"-L{libdir}", f"-l{lib}", f"-Wl,-rpath,{libdir}", "-lm", "-o", str(exe)]
    if width == 64:
        cmd.insert(1, "-DPATH_WIDTH=64")
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.split("\n")
    assert lines[0].split() == [str(width // 8), "32", "72", "80" if width == 32 else "152", "8"]
    assert lines[1].strip() == "1"
