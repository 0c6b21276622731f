"""Public-API throughput: Trainer.train() on an in-memory dataset of synthetic residue-level graphs (one B200).
Reports graphs/s of the training passes for the resident route (dataset collated once and kept in HBM; batches = id lists for the
per-graph step kernel, device-gathered Batches for every other network) and the streamed route (DRK_NO_RESIDENT=1: host collate of
every batch + PCIe copy, one batch ahead).
usage: python profiles/trainer_probe.py [graphs=2048] [ginet_nocluster|vanilla|ginet|foutnet|sgat]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeprank2_b200.dataset import InMemoryGraphDataset
from deeprank2_b200.neuralnets.gnn import foutnet, ginet, ginet_nocluster, sgat, vanilla_gnn
from deeprank2_b200.synthetic import make_graph
from deeprank2_b200.trainer import Trainer

n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
net_name = sys.argv[2] if len(sys.argv) > 2 else "ginet_nocluster"
GINet = {"ginet_nocluster": ginet_nocluster.GINet, "vanilla": vanilla_gnn.VanillaNetwork, "ginet": ginet.GINet, "foutnet": foutnet.FoutNet, "sgat": sgat.SGAT}[net_name]
clustered = net_name in ("ginet", "foutnet", "sgat")
graphs = [make_graph(g, with_clusters=clustered) for g in range(n_graphs)]
for mode in ("resident", "streamed"):
    if mode == "streamed":
        os.environ["DRK_NO_RESIDENT"] = "1"
    else:
        os.environ.pop("DRK_NO_RESIDENT", None)
    ds = InMemoryGraphDataset([g.clone() for g in graphs], clustering_method="mcl" if clustered else None)
    torch.manual_seed(0)
    trainer = Trainer(GINet, ds, cuda=True, output_exporters=[])
    trainer.train(nepoch=1, batch_size=256, validate=False, filename=None)  # warm-up: builds the resident set, compiles nothing
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    epochs = 3
    for e in range(epochs):
        trainer.model.train()
        trainer._epoch(e + 1, "training")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{net_name} {mode:9s}: {epochs * n_graphs / dt:10.0f} graphs/s through Trainer._epoch ({dt / epochs * 1e3:.1f} ms per epoch of {n_graphs} graphs, batch 256, loader {type(trainer.train_loader).__name__})", flush=True)
