"""Experiment: batch 8 as ONE launch plan vs two concurrent half-batch plans on two streams (tail / small-grid filling)."""
import sys, copy
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / 'lossy-vae_b200', ROOT / 'oracle'):
    sys.path.insert(0, str(p))
import lvae, lvae_oracle as O
from oracle_inputs import make_input
prec = sys.argv[1] if len(sys.argv) > 1 else 'f16x3+tail1'
nsplit = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = 8
models = []
for i in range(nsplit):
    m = lvae.get_model('qarv_base'); m.load_state_dict(O.sensitised_state_dict(O.qarv_param_shapes(), seed=0), strict=False)
    m.precision = prec
    models.append(m.cuda().eval())
im = make_input('synth', B, 512, 768, 1000).cuda()
def timeit(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
eng = models[0].engine
P = eng.forward_plan(B, 512, 768, 'eval'); P.im.copy_(im); P.lmb.fill_(2048.0)
t1 = timeit(lambda: eng.replay(P))
print(f'one plan, batch {B}: {t1:.3f} ms/step = {B / t1 * 1e3:.1f} images/s')
ref = P.stats.clone()
hb = B // nsplit
plans, streams = [], [torch.cuda.Stream() for _ in range(nsplit)]
for i, m in enumerate(models):
    Pi = m.engine.forward_plan(hb, 512, 768, 'eval'); Pi.im.copy_(im[i * hb:(i + 1) * hb]); Pi.lmb.fill_(2048.0)
    plans.append(Pi)
torch.cuda.synchronize()
for i, m in enumerate(models):           # warm + capture each on its own stream
    with torch.cuda.stream(streams[i]):
        for _ in range(3): m.engine.replay(plans[i])
torch.cuda.synchronize()
def dual():
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    for i, m in enumerate(models):
        streams[i].wait_event(ev)
        with torch.cuda.stream(streams[i]):
            m.engine.replay(plans[i])
        e2 = torch.cuda.Event(); e2.record(streams[i]); cur.wait_event(e2)
t2 = timeit(dual)
print(f'{nsplit} concurrent plans, batch {hb} each: {t2:.3f} ms/step = {B / t2 * 1e3:.1f} images/s')
th = timeit(lambda: models[0].engine.replay(plans[0]))
print(f'one plan, batch {hb} alone: {th:.3f} ms/step = {hb / th * 1e3:.1f} images/s')
