import faulthandler, os, sys, time
faulthandler.dump_traceback_later(45, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import petibm_b200 as pb
from petibm_b200.dist import Comm
from tests import helpers as H

def log(*a):
    print(f"[r{os.environ.get('RANK')}] ", *a, flush=True)

comm = Comm.from_env(reduce="nccl", halo=sys.argv[1] if len(sys.argv) > 1 else "memcpy")
torch.cuda.set_device(comm.device)
log("comm up")
shape, per = (24, 20, 44), (0, 0, 0)
widths = H.make_widths(shape)
s = pb.LinSolverB200("poisson", "None", comm=comm, device=comm.device)
log("solver created")
s.setOptions(rtol=0.0, atol=0.0, max_it=10)
if len(sys.argv) > 2: s.setTuning("use_graph", int(sys.argv[2]))
s.setStencil(H.grid_of(widths, per))
log("stencil set + connected")
s.setNullSpace(True)
xs = np.random.default_rng(0).standard_normal(int(np.prod(shape)))
xl = comm.local_block(xs, shape)
y = s.apply(xl)
log("apply done", float(np.abs(y).max()))
x = np.empty_like(xl)
try:
    s.solve(x, y)
except pb.B200Error as e:
    log("solve ended", e)
log("its", s.getIters(), "hist", s.getHistory()[-1])
s.destroy()
log("done")
