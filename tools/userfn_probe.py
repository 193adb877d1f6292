import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
osb = importlib.import_module("optimization-solvers_b200")
n = 96
cvec = torch.linspace(1.0, 4.0, n, dtype=torch.float64, device="cuda")
avec = torch.linspace(-0.5, 0.5, n, dtype=torch.float64, device="cuda")
ctx = osb.default_context()
ext = torch.cuda.ExternalStream(ctx.stream())
def T(ptr, count):
    class H: pass
    h = H(); h.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device="cuda")
calls = [0]
def enqueue(d_x, n_, d_f, d_g, d_h, stream):
    calls[0] += 1
    with torch.cuda.stream(ext):
        x = T(d_x, n_); g = T(d_g, n_); f = T(d_f, 1)
        dlt = x - avec
        g.copy_(cvec * dlt); f.copy_((0.5 * cvec * dlt * dlt).sum().reshape(1))
        if d_h:
            ld = (n_ + 15) // 16 * 16
            h = T(d_h, n_ * ld).view(n_, ld); h.zero_(); h[:, :n_].copy_(torch.diag(cvec))
    return 0
obj = osb.UserDeviceObjective(enqueue, n, with_hessian=True)
for cls, tol in (("GradientDescent", 1e-8), ("BFGS", 1e-8), ("Newton", 1e-10)):
    s = getattr(osb, cls)(tol, np.zeros(n)).record_trace(True)
    try:
        s.minimize(osb.BackTracking(1e-4, 0.5), obj, 60, 40); st = "Ok"
    except Exception as e:
        st = repr(e)
    tr = s.trace()
    print(cls, st, s.k(), s.termination_reason(), "f", tr["f"][:6], "t", tr["t"][:8], "err", np.max(np.abs(s.x() - avec.cpu().numpy())), "calls", calls[0])
