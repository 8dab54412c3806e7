"""Per-launch timeline of the dequant-GEMM launches inside a real (graph-replayed) decode step: for each launch the
earliest CTA start / latest CTA end (globaltimer, ns) and CTA 0's phase stamps.  Diagnostics only."""
import ctypes, sys, json
import torch
sys.path.insert(0, ".")
from exllamav2_b200 import ext as ext_c
from exllamav2_b200.model import ExLlamaV2Decoder, PRESETS

dev = "cuda:0"
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = PRESETS["llama2-7b-4.0bpw"]()
cfg.num_layers = layers
dec = ExLlamaV2Decoder(cfg, dev, seed=0, batch_size=1, cache_len=1024)
g = torch.Generator().manual_seed(0)
dec.prefill(torch.randint(0, cfg.vocab_size, (1, 128), generator=g).to(dev))
ext_c.lib.exl2b_debug_set.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
stamps = torch.zeros((64, 32), dtype=torch.int64, device=dev)
records = torch.zeros((64, 160, 4), dtype=torch.int64, device=dev)     # per CTA of every batch-1 GEMV launch: start, wait over, end, SM
ext_c.lib.exl2b_debug_set_records.argtypes = [ctypes.c_void_p]
dec._decode_step(); torch.cuda.synchronize()
ext_c.lib.exl2b_debug_set_records(records.data_ptr())
ext_c.lib.exl2b_debug_set(0, stamps.data_ptr(), 0)
dec.capture()                      # slots are assigned at capture: one per GEMM launch, in order
ext_c.lib.exl2b_debug_set(0, None, 0)
ext_c.lib.exl2b_debug_set_records(None)
ids = torch.zeros((1, 1), dtype=torch.long, device=dev)
for _ in range(3):
    dec.decode(ids)
torch.cuda.synchronize()
stamps.zero_(); stamps[:, 6] = 2**62
records.zero_()
dec.decode(ids); torch.cuda.synchronize()
st = stamps.cpu().tolist()
n = 5 * layers + 1
names = ["qkv", "attn", "o", "gateup", "down"]
t0 = st[0][6]
# the warm-up step in capture() used the first n slots; the captured graph the next n
rows = [r for r in st if r[6] < 2**62]
print("launches with stamps:", len(rows))
prev_end = None
for i, r in enumerate(rows):
    nm = names[i % 5] if i < 5 * layers else "head"
    gs, ge = r[6] - rows[0][6], r[7] - rows[0][6]
    c = [x - rows[0][6] if x else None for x in r[0:6]]
    print(f"{i:3d} {nm:7s} grid {gs/1e3:8.2f} -> {ge/1e3:8.2f} us  dur {(ge-gs)/1e3:6.2f}  gap_from_prev_end {((gs-prev_end)/1e3 if prev_end is not None else 0):6.2f}  cta0 "
          + " ".join("   -  " if x is None else f"{x/1e3:7.2f}" for x in c)
          + ("  [8,9] " + " ".join(f"{(x - rows[0][6])/1e3:7.2f}" for x in r[8:10] if x) if any(r[8:10]) else ""))
    prev_end = ge
print("step span us", (rows[-1][7] - rows[0][6]) / 1e3)
# per-CTA view of the GEMV launches: how many SMs host two CTAs of the SAME launch, spread of the CTAs' compute time
rec = records.cpu()
slot_of = [i for i, r in enumerate(st) if r[6] < 2**62]
from collections import Counter
for li, slot in enumerate(slot_of):
    r = rec[slot]
    live = r[:, 2] > 0
    if not bool(live.any()):
        continue
    rr = r[live]
    sm = Counter(rr[:, 3].tolist())
    dbl = sum(1 for v in sm.values() if v >= 2)
    comp = (rr[:, 2] - rr[:, 1]).float() / 1e3          # dependency wait over -> end
    late = (rr[:, 0].max() - rr[:, 0].min()).item() / 1e3
    nm = names[li % 5] if li < 5 * layers else "head"
    print(f"{li:3d} {nm:7s} ctas {int(live.sum()):4d} SMs {len(sm):4d} SMs-with-2 {dbl:3d}  wait->end us: min {comp.min():6.2f} med {comp.median():6.2f} max {comp.max():6.2f}"
          f"  start spread {late:7.2f}  last end - median end {(rr[:, 2].max() - rr[:, 2].median()).item() / 1e3:6.2f}")
