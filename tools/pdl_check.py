"""Single-GPU determinism check of the un-chained decode sequence (the one tensor_p / tools/tp_check.py compares against):
python tools/pdl_check.py <out.pt> [preset] -- run twice (with and without EXL2B_NO_PDL=1) and compare the saved logits."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exllamav2_b200.model import PRESETS, ExLlamaV2Decoder
out, preset = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "test-small")
dev = torch.device("cuda:0")
res = {}
for chained in (False, True):
    dec = ExLlamaV2Decoder(PRESETS[preset](), device=dev, seed=3, batch_size=1, cache_len=512)
    dec.chained = chained
    g = torch.Generator().manual_seed(5)
    prompt = torch.randint(0, dec.cfg.vocab_size, (1, 11), generator=g).to(dev)
    gen = torch.randint(0, dec.cfg.vocab_size, (1, 12), generator=g).to(dev)
    dec.prefill(prompt)
    res[chained] = torch.stack([dec.decode(gen[:, t:t + 1]).float().clone() for t in range(gen.shape[1])]).cpu()
    dec.unload()
torch.save(res, out)
if len(sys.argv) > 3:
    a = torch.load(sys.argv[3])
    for chained in (False, True):
        errs = [float((res[chained][t] - a[chained][t]).norm() / a[chained][t].norm()) for t in range(res[chained].shape[0])]
        print("chained" if chained else "unchained", "rel_l2 vs", sys.argv[3], " ".join(f"{e:.1e}" for e in errs))
    errs = [float((res[True][t] - res[False][t]).norm() / res[False][t].norm()) for t in range(res[True].shape[0])]
    print("chained vs unchained (this run)", " ".join(f"{e:.1e}" for e in errs))
