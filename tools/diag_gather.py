"""Timeline of one C3 step with the image all-gather on a side stream (2+ GPUs, torchrun): where does the time go?
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/diag_gather.py
Prints, per variant, the device time of the step and event offsets (ms from the start of the step) on rank 0.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, build_inputs  # noqa: E402
from jrender_b200 import SoftRasterizeFunction  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    nf, H, bpg, desc = WORKLOADS["c3"]
    fv_h, tex_h, grad_h = build_inputs("c3", rank, world)
    fv = torch.from_numpy(fv_h).to(dev).requires_grad_(True)
    tex = torch.from_numpy(tex_h).to(dev).requires_grad_(True)
    grad = torch.from_numpy(grad_h).to(dev)
    out = torch.empty((world * bpg, 4, H, H), device=dev)
    side = torch.cuda.Stream(dev)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    main_s = torch.cuda.current_stream(dev)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def variant(name):
        marks = {}

        def mark(k, stream=None):
            e = ev()
            e.record(stream or main_s)
            marks[k] = e
        fv.grad = None
        tex.grad = None
        flush.fill_(1.0)
        mark("t0")
        im = SoftRasterizeFunction(image_size=H)(fv, tex)
        mark("fwd_done")
        imd = im.detach()
        if name == "none":
            im.backward(grad)
        elif name == "blocking_before_bwd":
            dist.all_gather_into_tensor(out, imd)
            mark("gather_done")
            im.backward(grad)
        elif name == "blocking_after_bwd":
            im.backward(grad)
            mark("bwd_done")
            dist.all_gather_into_tensor(out, imd)
            mark("gather_done")
        elif name == "side_before_bwd":
            side.wait_stream(main_s)
            with torch.cuda.stream(side):
                mark("side_start", side)
                dist.all_gather_into_tensor(out, imd)
                mark("gather_done", side)
            im.backward(grad)
            mark("bwd_done")
            main_s.wait_stream(side)
        elif name == "side_after_bwd_launch":
            fwd_ev = torch.cuda.Event()
            fwd_ev.record(main_s)
            im.backward(grad)
            mark("bwd_done")
            side.wait_event(fwd_ev)
            with torch.cuda.stream(side):
                mark("side_start", side)
                dist.all_gather_into_tensor(out, imd)
                mark("gather_done", side)
            main_s.wait_stream(side)
        elif name == "async_op_before_bwd":
            w = dist.all_gather_into_tensor(out, imd, async_op=True)
            im.backward(grad)
            mark("bwd_done")
            w.wait()
        elif name == "side_gather_only":
            side.wait_stream(main_s)
            with torch.cuda.stream(side):
                mark("side_start", side)
                dist.all_gather_into_tensor(out, imd)
                mark("gather_done", side)
            main_s.wait_stream(side)
        mark("end")
        return marks

    names = ["none", "blocking_before_bwd", "blocking_after_bwd", "side_before_bwd", "side_after_bwd_launch",
             "async_op_before_bwd", "side_gather_only"]
    res = {}
    for name in names:
        for _ in range(3):
            variant(name)
        torch.cuda.synchronize()
        dist.barrier()
        acc = {}
        n = 8
        for _ in range(n):
            m = variant(name)
            torch.cuda.synchronize()
            for k, e in m.items():
                if k != "t0":
                    acc[k] = acc.get(k, 0.0) + m["t0"].elapsed_time(e) / n
        dist.barrier()
        # back-to-back, no host sync between steps (what bench.py times)
        a, b = ev(), ev()
        a.record()
        for _ in range(n):
            variant(name)
        b.record()
        torch.cuda.synchronize()
        acc["back_to_back_ms_per_step"] = a.elapsed_time(b) / n
        res[name] = {k: round(v, 3) for k, v in acc.items()}
        dist.barrier()
    if rank == 0:
        print(json.dumps(res, indent=1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
