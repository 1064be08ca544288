"""debug: does the caching allocator keep calling cudaMalloc in steady state?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rslo_b200
from rslo_b200.data import synthetic
from rslo_b200.utils.weights import deterministic_fill
from rslo_b200.utils.distributed import FlatGradAllReducer

dev = torch.device("cuda", 0)
net, vg = rslo_b200.build_network(testing=False, seed=7)
deterministic_fill(net, 11)
net = net.to(dev)
net.global_step.fill_(2000); net._step_host = None
net.train()
red = FlatGradAllReducer(net)
pairs = [synthetic.make_pair(s)[:2] for s in range(4)]
pairs = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b in pairs]

def step(i):
    red.zero_()
    pts = list(pairs[(2 * i) % 4]) + list(pairs[(2 * i + 1) % 4])
    ret = net({"points": pts, "n_samples": 2, "host_outputs": False})
    ret["loss"].sum().backward()
    return ret

for i in range(6):
    step(i)
torch.cuda.synchronize()
for i in range(6, 14):
    s0 = torch.cuda.memory_stats()
    t0 = time.time()
    ret = step(i)
    t1 = time.time()
    torch.cuda.synchronize()
    t2 = time.time()
    s1 = torch.cuda.memory_stats()
    print(f"step {i}: host {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms, cudaMalloc calls {s1['num_device_alloc']-s0['num_device_alloc']}, "
          f"cudaFree {s1['num_device_free']-s0['num_device_free']}, reserved {s1['reserved_bytes.all.current']/2**30:.2f} GiB, "
          f"active {s1['active_bytes.all.current']/2**30:.2f} GiB, alloc_retries {s1['num_alloc_retries']}")
    del ret

if os.environ.get("CPROFILE") == "1":
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for i in range(14, 24):
        step(i)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(45)
    st.sort_stats("cumulative").print_stats(60)
