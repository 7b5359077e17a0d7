import sys, pytest
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_train_srk_gpu as T
rec = []
orig = T.grad_close
def gc(got, want, name, rtol=1e-4):
    g, w = got.detach().cpu().double(), want.detach().cpu().double()
    scale = max(float(w.abs().max()), 1e-6); err = float((g - w).abs().max())
    rec.append((err / scale, name, tuple(w.shape)))
    return orig(got, want, name, rtol)
T.grad_close = gc
rc = pytest.main(["-q", "-m", "gpu", "-x", "/root/repo/tests/test_train_srk_gpu.py", "-k", "backward", "-p", "no:cacheprovider"])
rec.sort(reverse=True)
print("worst gradient ratios (err / max norm):")
for r in rec[:8]: print("  %.2e %s %s" % r)
print("n =", len(rec), "rc =", rc)
