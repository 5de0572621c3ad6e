"""The reference's own WAV tool, art.c, UNMODIFIED: compiled once on the reference's sources (oracle/_ref/art_ref, CPU) and once
against libresampler_b200.so (oracle/_ref/art_b200: resampler, biquad cascade and decimator all come from the library; only the
time stretcher stretch.c, which is not on the path, is the reference's own file).  Both are built by oracle/Makefile in the build
container and travel prebuilt.  Same WAV in, same options: the output files must have the same length, and -- integer PCM after
the float path -- may differ only where a float that differs by 1e-7 of peak rounds across an integer boundary: at most one LSB,
in a tiny fraction of the samples."""
import os
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REF = Path(__file__).resolve().parents[1] / "oracle" / "_ref"


def _write_wav(path, x, rate, bits=16):
    """x: (frames, ch) float in [-1, 1)"""
    ch = x.shape[1]
    if bits == 16:
        pcm = np.clip(np.round(x * 32767.0), -32768, 32767).astype("<i2").tobytes()
    else:
        v = np.clip(np.round(x * 8388607.0), -8388608, 8388607).astype("<i4")
        pcm = v.reshape(-1, 1).view(np.uint8)[:, :3].tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, ch, rate, rate * ch * bits // 8, ch * bits // 8, bits)
    path.write_bytes(hdr + b"data" + struct.pack("<I", len(pcm)) + pcm)


def _read_wav(path):
    b = path.read_bytes()
    at = 12
    fmt = None
    while at < len(b):
        cid, size = b[at:at + 4], struct.unpack("<I", b[at + 4:at + 8])[0]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", b[at + 8:at + 24])
        if cid == b"data":
            ch, bits = fmt[1], fmt[5]
            raw = np.frombuffer(b[at + 8:at + 8 + size], np.uint8)
            if bits == 16:
                return raw.view("<i2").reshape(-1, ch).astype(np.int64), fmt
            v = raw.reshape(-1, 3).astype(np.int64)
            val = v[:, 0] | (v[:, 1] << 8) | (v[:, 2] << 16)
            val = np.where(val >= 1 << 23, val - (1 << 24), val)
            return val.reshape(-1, ch), fmt
        at += 8 + size + (size & 1)
    raise AssertionError("no data chunk")


@pytest.mark.skipif(not (REF / "art_b200").exists() or not (REF / "art_ref").exists(), reason="oracle/_ref/art_* were not built (no reference sources at build time)")
@pytest.mark.parametrize("name,src,ch,bits,opts", [
    ("up", 44100, 2, 16, ["-3", "-r48000", "-d0", "-n0", "-x"]),                      # BASELINE config 2 as art runs it (fixed ratio, 160 filters)
    ("down_prefilter", 96000, 2, 24, ["-3", "-r44100", "-p", "-d0", "-n0", "-x"]),   # downsampling with the biquad pre-filter cascade (-p)
    ("default", 44100, 2, 16, ["-2", "-r32k"]),                                        # art's defaults: endpoint extrapolation, HP tpdf dither, ATH shaping
    ("mono_up_post", 22050, 1, 16, ["-1", "-r48000", "-p", "-d1", "-n2"]),             # upsampling with the post-filter cascade, flat dither, 2nd-order shaping
])
def test_art_tool_on_the_library_matches_art_on_the_reference(tmp_path, name, src, ch, bits, opts):
    rng = np.random.default_rng(len(name))
    n = src * 3
    t = np.arange(n)[:, None] / src
    x = 0.25 * np.sin(2 * np.pi * (440.0 + 110.0 * np.arange(ch)) * t) + 0.15 * np.sin(2 * np.pi * 5000.0 * t) + rng.uniform(-0.2, 0.2, (n, ch))
    win = tmp_path / "in.wav"
    _write_wav(win, x, src, bits)
    outs = {}
    for which in ("ref", "b200"):
        out = tmp_path / f"out_{which}.wav"
        r = subprocess.run([str(REF / f"art_{which}"), "-q", "-y", *opts, str(win), str(out)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (which, r.stdout[-500:], r.stderr[-500:])
        outs[which] = _read_wav(out)
    (a, fa), (b, fb) = outs["ref"], outs["b200"]
    assert fa == fb and a.shape == b.shape, "header or length differs"
    d = np.abs(a - b)
    dithered = "-d0" not in opts
    full = float(1 << (fa[5] - 1))
    if not dithered:
        # the float paths agree within 1e-6 of full scale: one LSB at 16 bits (and then only where a sample sits on a rounding
        # boundary), a handful of LSBs at 24 bits, whose LSB (1.2e-7) is finer than float arithmetic itself
        assert d.max() <= max(1, int(1e-6 * full) + 1), (int(d.max()), float(np.mean(d > 0)))
        if fa[5] <= 16:
            assert np.mean(d > 0) <= 2e-3, float(np.mean(d > 0))
    else:
        # with noise shaping the quantiser error is fed back: one flipped rounding decision (a float differing by 1e-7) changes
        # the shaper's state and the following decisions, so the two files are two equally valid dithered renderings of the same
        # signal -- they must agree to within the shaped noise itself (a few LSB peak, about one LSB rms); the decimator's own
        # bit-exactness on identical floats is what tests/test_gpu_decimator.py pins
        rms = float(np.sqrt(np.mean(d.astype(np.float64) ** 2)))
        assert d.max() <= 64 and rms <= 2.0, (int(d.max()), rms)
