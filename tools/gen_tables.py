#!/usr/bin/env python
"""Prints the float literal tables shared (by value, not by file) between oracle/ and csrc/.
WcMul[n][i] = 1/(2 cos((i+1/2) pi / n)); Resample[n_from->n_to][k] = sin(pi k/2M)/((N/M) sin(pi k/2N))."""
import numpy as np
def lit(v):
    return repr(float(np.float32(v))) + "f" if "e" in repr(float(np.float32(v))) or "." in repr(float(np.float32(v))) else repr(float(np.float32(v))) + ".0f"
def f(v):
    s = np.format_float_scientific(np.float32(v), unique=True)
    return s + "f"
for n in (4, 8, 16, 32):
    vals = [1.0 / (2.0 * np.cos((i + 0.5) * np.pi / n)) for i in range(n // 2)]
    print(f"static const float kWc{n}[{n//2}] = {{" + ", ".join(f(v) for v in vals) + "};")
for (N, M) in ((16, 2), (32, 4)):
    vals = [1.0 if k == 0 else np.sin(np.pi * k / (2.0 * M)) / ((N // M) * np.sin(np.pi * k / (2.0 * N))) for k in range(M)]
    print(f"static const float kResample{N}_{M}[{M}] = {{" + ", ".join(f(v) for v in vals) + "};")
# 64-point transforms (round 2): 1/(2 cos((i+1/2) pi / 64)) and the 64 -> 8 resample scales
vals = [1.0 / (2.0 * np.cos((i + 0.5) * np.pi / 64)) for i in range(32)]
print("static const float kWc64[32] = {" + ", ".join(f(v) for v in vals) + "};")
vals = [1.0 if k == 0 else np.sin(np.pi * k / 16.0) / (8 * np.sin(np.pi * k / 128.0)) for k in range(8)]
print("static const float kResample64_8[8] = {" + ", ".join(f(v) for v in vals) + "};")
