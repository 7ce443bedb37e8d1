#!/usr/bin/env python
"""Does this torch build capture a torch.distributed all_reduce into a CUDA graph (a) on the capture stream, (b) on a side
stream forked with events (what engine.Program.run_eager does for lanes)?  Run under torchrun with 2 ranks."""
import os, sys, torch, torch.distributed as dist
variant = sys.argv[1]
rank = int(os.environ['RANK']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl')
t = torch.full((1 << 20,), float(rank + 1), device='cuda')
dist.all_reduce(t)                        # communicator set-up outside the capture
torch.cuda.synchronize()
print('rank %d eager all_reduce ok: %.1f' % (rank, float(t[0])), flush=True)
t.fill_(float(rank + 1))
side = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    t.mul_(2.0)
    if variant == 'main':
        dist.all_reduce(t)
    else:
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
        with torch.cuda.stream(torch.cuda.ExternalStream(side.cuda_stream)):
            dist.all_reduce(t)
        ev2 = torch.cuda.Event(); ev2.record(side); main.wait_event(ev2)
    t.add_(1.0)
print('rank %d captured (%s)' % (rank, variant), flush=True)
for i in range(3):
    t.fill_(float(rank + 1))
    g.replay()
    torch.cuda.synchronize()
    print('rank %d replay %d: %.1f (expect 7.0)' % (rank, i, float(t[0])), flush=True)
dist.destroy_process_group()
