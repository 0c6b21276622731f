"""Where the host time of Trainer._epoch goes: python profiles/trainer_cprofile.py [graphs=1024] [net=sgat]"""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeprank2_b200.dataset import InMemoryGraphDataset
from deeprank2_b200.neuralnets.gnn import foutnet, ginet, ginet_nocluster, sgat, vanilla_gnn
from deeprank2_b200.synthetic import make_graph
from deeprank2_b200.trainer import Trainer

n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net_name = sys.argv[2] if len(sys.argv) > 2 else "sgat"
Net = {"ginet_nocluster": ginet_nocluster.GINet, "vanilla": vanilla_gnn.VanillaNetwork, "ginet": ginet.GINet, "foutnet": foutnet.FoutNet, "sgat": sgat.SGAT}[net_name]
clustered = net_name in ("ginet", "foutnet", "sgat")
ds = InMemoryGraphDataset([make_graph(g, with_clusters=clustered) for g in range(n_graphs)], clustering_method="mcl" if clustered else None)
torch.manual_seed(0)
trainer = Trainer(Net, ds, cuda=True, output_exporters=[])
trainer.train(nepoch=1, batch_size=256, validate=False, filename=None)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for e in range(3):
    trainer.model.train()
    trainer._epoch(e + 1, "training")
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
