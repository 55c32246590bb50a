"""Experiment: tcgen05 attention context error for candidate MN-major V descriptor strides (PFASR_ATT_VDESC=lbo,sbo)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from aliparaformerasr_b200 import _lib
from oracle import sanm
from _util import half_round
lib = _lib.load(); rng = np.random.default_rng(7)
B, Tq, Tk, H, D = 2, %d, %d, 4, 512
q = rng.standard_normal((B, Tq, D)).astype(np.float32); k = rng.standard_normal((B, Tk, D)).astype(np.float32); v = rng.standard_normal((B, Tk, D)).astype(np.float32)
out = np.zeros_like(q)
_lib.check(lib.pf_dbg_attention(B, H, Tq, Tk, _lib.fptr(q), _lib.fptr(k), _lib.fptr(v), _lib.fptr(out)))
ref = sanm._mha(torch.from_numpy(half_round(q)), torch.from_numpy(half_round(k)), torch.from_numpy(half_round(v)), H).numpy()
print('max err', float(np.abs(out - ref).max()), 'nan', int(np.isnan(out).sum()))
"""
for T in (64, 166):
    tkp = (T + 15) // 16 * 16
    box = tkp * 128
    for lbo, sbo in [(box, 1024), (1024, box), (box, 2048), (2048, box), (128, 1024), (1024, 128), (box, 128), (16, 1024), (1024, 16)]:
        env = dict(os.environ, PFASR_ATT_VDESC=f"{lbo},{sbo}")
        r = subprocess.run([sys.executable, "-c", CODE % (ROOT, ROOT, T, T)], env=env, capture_output=True, text=True, timeout=120)
        print(f"T={T} lbo={lbo} sbo={sbo}:", (r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
