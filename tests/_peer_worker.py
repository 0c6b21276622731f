"""Worker of tests/test_gpu_multi.py (run under torchrun, one process per GPU): data-parallel GINet steps whose gradient
all-reduce happens inside the finalize kernel over NVLink peer memory, checked against a single-process run on the union of the
ranks' graphs, for bit-identical weights on all ranks, and with a rank that has no graphs in a ragged step."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeprank2_b200.data import Batch  # noqa: E402
from deeprank2_b200.fused import GINetFusedStep  # noqa: E402
from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet  # noqa: E402
from deeprank2_b200.synthetic import make_graph  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    per = 10
    graphs = [make_graph(g) for g in range(per * world)]
    mine = Batch.from_data_list([g.clone() for g in graphs[rank * per : (rank + 1) * per]]).to(dev)
    everything = Batch.from_data_list([g.clone() for g in graphs]).to(dev)

    def build(world_size):
        torch.manual_seed(0)
        net = GINet(50, 1, 1).to(dev).eval()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
        return net, opt, GINetFusedStep(net, opt, torch.nn.MSELoss(), world_size=world_size)

    net, opt, step = build(world)
    mode = "peer" if step._peers is not None else "nccl"
    if os.environ.get("DRK_REQUIRE_PEER") and mode != "peer":
        raise RuntimeError("the peer-memory exchange was not set up")
    ref_net, ref_opt, ref_step = build(1)

    def assert_grads_match(what):
        g, r = step.flat_grad, ref_step.flat_grad
        bound = 1e-5 * float(r.abs().max()) + 1e-5 * r.abs()
        worst = float(((g - r).abs() - bound).max())
        assert worst <= 0.0, f"{what}: data-parallel gradient differs from the single-process gradient (excess {worst:.3e}, max|g| {float(r.abs().max()):.3e})"

    # the exchanged gradient itself, before any optimizer arithmetic: the sum over ranks must be the single-process gradient (fp32 bar)
    step.forward_backward(mine, global_size=per * world)
    ref_step.forward_backward(everything)
    torch.cuda.synchronize()
    assert_grads_match(f"{mode} exchange")
    losses = []
    for i in range(3):
        if mode == "nccl":
            # torch's own Adam applies this mode's update: compare GRADIENTS from identical weights at every step instead of weights
            # after sign-like Adam steps (a near-zero gradient component may legitimately move its weight by lr either way)
            net.load_state_dict(ref_net.state_dict())
        loss, _ = step(mine, global_size=per * world)
        ref_loss, _ = ref_step(everything)
        if mode == "nccl":
            torch.cuda.synchronize()
            assert_grads_match(f"nccl step {i}")
        total = loss.clone()
        if mode == "nccl":
            dist.all_reduce(total)  # the NCCL path leaves the rank-local share in `loss`; the peer path already holds the global value
        losses.append((float(total), float(ref_loss)))
    # ragged step: only rank 0 has graphs
    if rank == 0:
        step(mine, global_size=per)
    else:
        step.empty_step()
    ref_step(Batch.from_data_list([g.clone() for g in graphs[:per]]).to(dev))
    torch.cuda.synchronize()
    # peer mode: same kernels, same Adam arithmetic, gradients equal up to the association of the cross-rank sum -> tight bars.
    # nccl mode: the update is torch's own Adam; its first steps are sign-like (dw ~ lr * g / |g|), so gradients that agree to
    # 1e-7 can still move a near-zero-gradient weight by a fraction of lr: the first loss is tight, the rest within lr-sized drift.
    tight = mode == "peer"
    for i, (a, b) in enumerate(losses):
        assert abs(a - b) <= 1e-5 * abs(b) + 1e-7, f"loss {a} vs single-process {b} at step {i}"
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    ref = torch.cat([p.detach().reshape(-1) for p in ref_net.parameters()])
    diff = (flat - ref).abs()
    err = float(diff.max())
    if tight:
        assert err <= 2e-6, f"weights after 4 data-parallel steps differ from the single-process run by {err:.3e}"
    # (nccl mode: gradients were compared step by step above; the weights are only required to be identical across ranks)
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], t) for t in gathered), "ranks hold different weights"
    for p in net.parameters():
        assert float(opt.state[p]["step"]) == 4.0
    print(f"PEER_OK rank {rank} mode {mode} max|dw| {err:.2e}", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
