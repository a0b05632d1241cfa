"""torch.profiler (CUPTI) view of ONE eager training step: GPU time per phase and per kernel, split into
library (sr::) kernels and everything torch launches on our behalf (the fusion to-do list)."""
import collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile, record_function
import bench
from sradsgan_b200.model.sradsgan import SRADSGAN

B = int(os.environ.get("SR_BATCH", "16"))
net = SRADSGAN(bench.trainer_args(batch_size=B))
net.build(init=True)
hr = torch.rand(B, 3, 216, 216, device="cuda")
lr = torch.nn.functional.interpolate(hr, size=54, mode="bicubic", align_corners=False).clamp(0, 1)
for _ in range(2):
    net.train_step(lr, hr)
torch.cuda.synchronize()
ctx = {"cur": None}
def mark(name):
    if ctx["cur"] is not None:
        ctx["cur"].__exit__(None, None, None)
    ctx["cur"] = record_function("PHASE:" + name)
    ctx["cur"].__enter__()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    net._phase_mark = mark
    net.train_step(lr, hr)
    torch.cuda.synchronize()
    if ctx["cur"] is not None:
        ctx["cur"].__exit__(None, None, None)
net._phase_mark = None
evs = prof.events()
phases = sorted([(e.time_range.start, e.time_range.end, e.name[6:]) for e in evs if e.name.startswith("PHASE:")])
# phase X covers the CPU interval AFTER mark(X) until the next mark, i.e. the work of the NEXT phase name
names = [p[2] for p in phases]
def phase_of(t):
    for (s, e, n), nxt in zip(phases, names[1:] + ["end"]):
        if s <= t < e:
            return nxt
    return "?"
# map kernels to phases through their launching CPU op (correlation by launch time)
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0.0]))
for e in evs:
    if e.device_type == torch.autograd.DeviceType.CUDA:
        continue
    for k in e.kernels:
        short = re.sub(r"<.*", "", k.name).replace("void ", "")[:60]
        if not short.startswith("sr::"):
            shp = ""
            try:
                shp = str([tuple(x) for x in e.input_shapes if x][:2])
            except Exception:
                pass
            short = (short.replace("at::native::", "")[:28] + " <- " + e.name.replace("aten::", "") + " " + shp)[:100]
        a = agg[phase_of(e.time_range.start)][short]
        a[0] += 1; a[1] += k.duration
seen = set()
for ph in names[1:] + ["end", "?"]:
    if ph in seen or ph not in agg:
        continue
    seen.add(ph)
    tot = sum(v[1] for v in agg[ph].values())
    ours = sum(v[1] for k, v in agg[ph].items() if k.startswith("sr::"))
    print("== %-18s total %8.2f ms | sr:: %8.2f ms | other %8.2f ms | launches %d" % (ph, tot / 1e3, ours / 1e3, (tot - ours) / 1e3, sum(v[0] for v in agg[ph].values())))
    for k, v in sorted(agg[ph].items(), key=lambda kv: -kv[1][1])[:24]:
        print("      %-100s %5d %9.3f ms" % (k, v[0], v[1] / 1e3))
# whole-step table of everything that is NOT a library kernel, by (op, shapes)
tot = collections.defaultdict(lambda: [0, 0.0])
for ph in agg:
    for k, v in agg[ph].items():
        if not k.startswith("sr::"):
            key = ph + " | " + k
            tot[key][0] += v[0]; tot[key][1] += v[1]
print("== non-library kernels of the whole step, by phase/op/shape (top 90 by launches)")
for k, v in sorted(tot.items(), key=lambda kv: (-kv[1][0], -kv[1][1]))[:90]:
    print("      %-125s %5d %9.3f ms" % (k[:125], v[0], v[1] / 1e3))
