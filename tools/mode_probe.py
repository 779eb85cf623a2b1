"""Single-GPU probe: cost of a fused CX_TSP + two-map op by shared-memory access mode (64-bit blocks when
tile digit 0 is free, 128-bit pairs when it is an op digit).  Prints ms for 8 / 4 ops and the per-op slope."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qiskit_aakash_b200 import capi, engine as eng, schedule
n_bits = 28
alloc = eng.TorchCudaAllocator(0)
ctx = eng.shared_context(capi.load_library(), 0)
ctx.set_stream(alloc.stream())
state = alloc.empty(1 << n_bits); state.fill_(0.5)
rng = np.random.default_rng(5)
def make_pass(td, pairs, kind):
    P = np.zeros(1, dtype=capi.PASS_DTYPE); K=len(td)
    P[0]["n_tile_digits"]=K; P[0]["tile_digit"][:K]=td
    for k,(la,lb) in enumerate(pairs):
        op=P[0]["ops"][k]; op["kind"]=kind; op["a"],op["b"]=la,lb
        if kind != capi.OP_SWAP:
            op["flags"]=3
            m=np.linalg.qr(rng.normal(size=(3,3)))[0]; full=np.zeros((3,4)); full[:,1:]=m
            op["pa"]=full.ravel(); op["pb"]=full.ravel(); op["coef"][:5]=eng.cx_coefficients((0.999,0.0))
        op["fd"][:K-2]=schedule.lane_order(K,la,lb)
    # explicit non-trailing guard: add a final MATS-free CX so swaps are not folded
    P[0]["n_ops"]=len(pairs); return P
def timed(P, reps=10):
    for _ in range(3): ctx.apply_passes(state.data_ptr(), n_bits, P)
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ctx.apply_passes(state.data_ptr(), n_bits, P)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
td=[0,1,5,6,7,8]
sets={"modeA (2,3)(4,5)":[(2,3),(4,5)]*4, "modeA (1,2)(3,4)":[(1,2),(3,4)]*4, "pairA (0,2)(0,3)":[(0,2),(0,3),(0,4),(0,5)]*2,
      "pairB (2,0)(3,0)":[(2,0),(3,0),(4,0),(5,0)]*2, "mixed (0,2)(3,4)":[(0,2),(3,4),(0,5),(1,3)]*2}
for name,pairs in sets.items():
    for kind,kn in ((capi.OP_CX_TSP,"tsp+mats"),):
        t8=timed(make_pass(td,pairs,kind)); t4=timed(make_pass(td,pairs[:4],kind))
        print(json.dumps({"set":name,"kind":kn,"ms_8ops":round(t8,4),"ms_4ops":round(t4,4),"per_op":round((t8-t4)/4,4)}))
