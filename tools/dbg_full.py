import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from tests.parity import *
for backend in ("cuda","jit"):
    c = gradient_case(backend, "quadcopter", B=64, T=100)
    r_ref = c["ref"]["reward"].detach().numpy().T
    print(backend, "reward rel", np.abs(c["got"]["reward"]-r_ref).max()/np.abs(r_ref).max())
    for s in c["model"].state_fluents:
        a = c["ref"]["final"][s].detach().reshape(64,-1).numpy(); b = c["final"][s].reshape(64,-1).numpy()
        print("  ", s, np.abs(a-b).max(), np.abs(a).max())
    g,gr=c["grad"],c["grad_ref"]
    print("  grad maxabs", np.abs(g-gr).max(), "gradmax", np.abs(gr).max())
