"""2-GPU check: DistributedDataParallel(model) in the reference's loop shape, eager vs AutoGraphedTrain.
torchrun --nproc-per-node 2 scripts/ddp_autograph_check.py"""
import os, sys, time
from pathlib import Path
import torch
import torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / 'lossy-vae_b200', ROOT / 'oracle'):
    sys.path.insert(0, str(p))
import lvae
from oracle_inputs import make_input
rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
res = {}
for mode in ('eager', 'autograph'):
    os.environ['LVAE_TRAIN_AUTOGRAPH_DDP'] = '1' if mode == 'autograph' else '0'
    torch.manual_seed(0)
    model = lvae.get_model('qarv_base').to(dev).train()
    model.train_path.autograph_enabled = mode == 'autograph'
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    ims = [make_input('synth', 16, 256, 256, 100 * rank + i).to(dev) for i in range(4)]
    lmb = torch.linspace(32, 2048, 16, device=dev)
    losses = []
    torch.manual_seed(7 + rank)
    for it in range(14):
        if it == 4:
            torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        out = ddp(ims[it % 4], lmb=lmb)
        opt.zero_grad(set_to_none=True)
        out['loss'].backward()
        opt.step()
        losses.append(out['loss'].item())
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / 10
    chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum()
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    res[mode] = (losses, dt, [float(c) for c in allc], model.train_path.autograph.core is not None)
    del ddp, model, opt
    torch.cuda.empty_cache()
if rank == 0:
    for mode, (losses, dt, allc, used) in res.items():
        print(f'{mode}: graphs used {used}; {16 * world / dt:.1f} images/s ({dt * 1e3:.1f} ms per step); losses {losses[0]:.3f} -> {losses[-1]:.3f}; '
              f'parameter checksums equal across ranks: {abs(allc[0] - allc[1]) <= 1e-9 * abs(allc[0])}')
    le, la = res['eager'][0], res['autograph'][0]
    print('max relative loss difference eager vs autograph over 14 steps:', max(abs(a - b) / abs(a) for a, b in zip(le, la)))
dist.destroy_process_group()
