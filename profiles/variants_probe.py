"""Config C4: train-step time of the other networks of the path (layer kernels through autograd) on the C2 batch, eager vs CUDA-graph replay."""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeprank2_b200 import _lib
from deeprank2_b200.neuralnets.gnn import foutnet, ginet, ginet_nocluster, sgat, vanilla_gnn
from deeprank2_b200.step import GraphedTrainStep, TrainStep
from deeprank2_b200.synthetic import make_batch

dev = torch.device("cuda", 0)
plain = make_batch(256).to(dev)
clustered = make_batch(256, with_clusters=True).to(dev)
loss_fn = torch.nn.MSELoss()
# the clustered networks size their pooled batch on the host (as the reference does): eager only
for name, cls, batch, capturable in (("VanillaNetwork", vanilla_gnn.VanillaNetwork, plain, True), ("GINet no-cluster, layer kernels", ginet_nocluster.GINet, plain, True),
                                     ("FoutNet (clustered)", foutnet.FoutNet, clustered, False), ("GINet (clustered)", ginet.GINet, clustered, False),
                                     ("SGAT (clustered)", sgat.SGAT, clustered, False)):
    if len(sys.argv) > 1 and sys.argv[1].lower() not in name.lower():
        continue
    torch.manual_seed(0)
    net = cls(50, 1, 1).to(dev).train()
    if hasattr(net, "fused"):
        net.fused = False
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
    inner = TrainStep(net, opt, loss_fn)

    def step(b, inner=inner):
        # FoutNet / clustered GINet / SGAT overwrite data.x and pool the batch in place, exactly like the reference (foutnet.py:104):
        # hand every step a fresh shallow view, as a loader would
        view = copy.copy(b)
        view.__dict__ = dict(b.__dict__)
        return inner(view)

    for _ in range(3):
        step(batch)
    torch.cuda.synchronize()
    c0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(20):
        step(batch)
    torch.cuda.synchronize()
    eager = (time.perf_counter() - t0) / 20
    launches = (_lib.launch_count() - c0) / 20
    try:
        if not capturable:
            raise RuntimeError("host-sized pooling")
        g = GraphedTrainStep(step, batch, warmup=1)
        for _ in range(3):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        graphed = a.elapsed_time(b) / 50
    except Exception as exc:  # noqa: BLE001
        graphed = float("nan")
        if capturable:
            print("   graph capture failed:", type(exc).__name__, str(exc)[:120])
    print(f"{name:34s} eager {eager * 1e3:7.3f} ms/step  graph replay {graphed:7.3f} ms/step  ({256 / graphed * 1e3:9.0f} graphs/s)  {launches:.0f} drk launches/step", flush=True)
