"""Derives the encoder-side 5x5 kernel that approximately inverts the decoder's default Gaborish blur (ISO/IEC 18181-1
loop filter, default weights: the 3x3 kernel [w2 w1 w2; w1 1 w1; w2 w1 w2] / (1 + 4 w1 + 4 w2) with w1 = 0.115169525,
w2 = 0.061248592).  libjxl ships hand-tuned constants for this (enc_gaborish.cc), which are not available offline; any
kernel is a legal encoder choice.  This one is the least-squares solution: the symmetric 5x5 kernel K (six distinct
weights) that minimises || K * G - delta ||^2 over the 7x7 support, renormalised to unit sum so that flat areas are
preserved exactly.  Prints the six weights as float literals for oracle/jxo_xyb.cc and csrc/k_gab.cu."""
import numpy as np

w1, w2 = 0.115169525, 0.061248592
G = np.array([[w2, w1, w2], [w1, 1.0, w1], [w2, w1, w2]], dtype=np.float64)
G /= G.sum()
classes = {}
for y in range(-2, 3):
    for x in range(-2, 3):
        classes.setdefault(tuple(sorted((abs(x), abs(y)), reverse=True)), []).append((y, x))
keys = sorted(classes)            # (0,0) (1,0) (1,1) (2,0) (2,1) (2,2)
A = np.zeros((49, len(keys)))
for j, k in enumerate(keys):
    K = np.zeros((5, 5))
    for (y, x) in classes[k]:
        K[y + 2, x + 2] = 1.0
    full = np.zeros((7, 7))
    for y in range(5):
        for x in range(5):
            full[y:y + 3, x:x + 3] += K[y, x] * G
    A[:, j] = full.ravel()
delta = np.zeros((7, 7)); delta[3, 3] = 1.0
sol, *_ = np.linalg.lstsq(A, delta.ravel(), rcond=None)
total = sum(sol[j] * len(classes[k]) for j, k in enumerate(keys))
sol = sol / total
res = A @ sol - delta.ravel()
print("classes (|dx|,|dy|) sorted:", keys)
print("weights:", ", ".join(f"{np.float32(v):.9e}f" for v in sol))
print("sum:", sum(sol[j] * len(classes[k]) for j, k in enumerate(keys)), "residual rms:", np.sqrt((res ** 2).mean()))
