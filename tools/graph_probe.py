"""CUDA-graph replay of the whole ClipLoss step (forward + mean + backward through the drop-in module) at the shapes the
FLYP loop actually runs: the library's calls allocate nothing, never synchronise and keep no host state for world_size 1,
so a training step that contains them can be captured; this measures what is left when the framework's per-call host
cost is out of the way, next to the eager reference operators captured the same way.
    python tools/graph_probe.py [B,B,...] [D]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F
import flyp_b200
import _inputs as torch_port

Bs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [256, 512, 4096]
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")


def eager_loss(I, T, s):
    li = s * I @ T.T
    lt = s * T @ I.T
    lab = torch.arange(li.shape[0], device=I.device, dtype=torch.long)
    return (F.cross_entropy(li, lab, reduction='none') + F.cross_entropy(lt, lab, reduction='none')) / 2


for B in Bs:
    for dt in (torch.bfloat16, torch.float32):
        I, T = torch_port.synthetic_pairs(B, D, seed=0, dtype=dt)
        out = {"B": B, "D": D, "dtype": str(dt).split(".")[-1]}
        for name, fn in (("flyp_b200", flyp_b200.ClipLoss(cache_labels=True)), ("eager", eager_loss)):
            Id = I.to(dev).requires_grad_(True); Td = T.to(dev).requires_grad_(True)
            theta = torch.tensor(2.6592600369327783, device=dev, requires_grad=True)

            def step():
                loss = fn(Id, Td, theta.exp())
                loss.mean().backward()
                return loss

            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    Id.grad = Td.grad = theta.grad = None
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            try:
                graph = torch.cuda.CUDAGraph()
                Id.grad = Td.grad = theta.grad = None
                with torch.cuda.graph(graph):
                    loss = step()
                for _ in range(5):
                    graph.replay()
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(100):
                    graph.replay()
                e1.record(); torch.cuda.synchronize()
                out[name + "_graph_us"] = round(e0.elapsed_time(e1) * 10, 1)
                out[name + "_loss0"] = round(float(loss.float().mean()), 5)
            except Exception as exc:  # noqa: BLE001
                out[name + "_graph_us"] = f"capture failed: {type(exc).__name__}: {str(exc)[:80]}"
            del graph
        # torch.cuda.make_graphed_callables: what a training loop can use as it is (forward and backward captured as two
        # graphs behind an autograd function, static input / output buffers managed by torch)
        try:
            Id = I.to(dev).requires_grad_(True); Td = T.to(dev).requires_grad_(True)
            sc = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
            mod = flyp_b200.ClipLoss(cache_labels=True)
            gfn = torch.cuda.make_graphed_callables(mod, (Id, Td, sc))
            want = mod(Id, Td, sc).float().mean().item()

            def gstep():
                Id.grad = Td.grad = sc.grad = None
                gfn(Id, Td, sc).mean().backward()

            for _ in range(5):
                gstep()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100):
                gstep()
            e1.record(); torch.cuda.synchronize()
            out["flyp_b200_graphed_callable_us"] = round(e0.elapsed_time(e1) * 10, 1)
            out["graphed_callable_loss_matches"] = abs(gfn(Id, Td, sc).float().mean().item() - want) < 1e-2 * abs(want)
        except Exception as exc:  # noqa: BLE001
            out["flyp_b200_graphed_callable_us"] = f"failed: {type(exc).__name__}: {str(exc)[:120]}"
        print(json.dumps(out), flush=True)
