"""Philox4x32-10 of include/ux_rng.h against the Random123 known-answer vectors, and the derived uniforms."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <stdio.h>
#include "ux_rng.h"
int main(void) {
    ux_u32x4 r = ux_philox4x32_10(0, 0, 0, 0, 0, 0);
    printf("%08x %08x %08x %08x\n", r.v[0], r.v[1], r.v[2], r.v[3]);
    r = ux_philox4x32_10(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    printf("%08x %08x %08x %08x\n", r.v[0], r.v[1], r.v[2], r.v[3]);
    r = ux_philox4x32_10(0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u, 0xa4093822u, 0x299f31d0u);
    printf("%08x %08x %08x %08x\n", r.v[0], r.v[1], r.v[2], r.v[3]);
    double s = 0; int hist[4] = {0, 0, 0, 0};
    for (int i = 0; i < 100000; i++) { double u = ux_rng_uniform(7, i, 3, UX_RNG_TRANSFER, 1); if (u < 0 || u >= 1) return 2; s += u; hist[ux_rng_below(7, i, 3, UX_RNG_GENERATE, 0, 4)]++; }
    printf("%.6f %d %d %d %d\n", s / 100000, hist[0], hist[1], hist[2], hist[3]);
    return 0;
}
'''


def test_philox_known_answers(tmp_path):
    c = tmp_path / "kat.c"
    c.write_text(SRC)
    exe = tmp_path / "kat"
    subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(c)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0] == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"   # Random123 kat_vectors: philox4x32 10 rounds, zeros
    assert out[1] == "408f276d 41c83b0e a20bc7c6 6d5451fd"   # all ones
    assert out[2] == "d16cfe09 94fdcceb 5001e420 24126ea1"   # pi digits
    mean, *hist = out[3].split()
    assert abs(float(mean) - 0.5) < 0.005
    assert all(abs(int(h) - 25000) < 800 for h in hist)
