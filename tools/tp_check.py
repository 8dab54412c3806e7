"""torchrun --nproc-per-node N tools/tp_check.py [preset]: sharded decode (exllamav2_b200/tensor_p.py) against the single-GPU
decoder on rank 0 -- same synthetic weights, same token ids; prints the relative L2 error of the logits per step."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exllamav2_b200.model import PRESETS, ExLlamaV2Decoder
from exllamav2_b200.tensor_p import ExLlamaV2DecoderTP

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
preset = sys.argv[1] if len(sys.argv) > 1 else "test-small"
cfg = PRESETS[preset]()
tpd = ExLlamaV2DecoderTP(cfg, rank, world, dev, seed=3, batch_size=1, cache_len=512)
g = torch.Generator().manual_seed(5)
prompt = torch.randint(0, cfg.vocab_size, (1, 11), generator=g).to(dev)
gen = torch.randint(0, cfg.vocab_size, (1, 4), generator=g).to(dev)
tpd.prefill(prompt)
outs = [tpd.decode(gen[:, t:t + 1]).float().clone() for t in range(gen.shape[1])]
graph_ok = True
try:
    tpd.capture()
    outs_g = [tpd.decode(gen[:, t:t + 1]).float().clone() for t in range(2)]
except Exception as e:
    graph_ok = False
    print(f"rank {rank}: graph capture failed: {e}")
ok = True
if rank == 0:
    ref = ExLlamaV2Decoder(PRESETS[preset](), device=dev, seed=3, batch_size=1, cache_len=512)
    ref.chained = False
    ref.prefill(prompt)
    for t in range(gen.shape[1]):
        r = ref.decode(gen[:, t:t + 1]).float()
        err = float((outs[t] - r).norm() / r.norm())
        print(f"step {t}: rel_l2(tp, single) = {err:.3e} finite={bool(torch.isfinite(outs[t]).all())}")
        ok &= err < 1e-2
    print("TP_CHECK", "PASS" if ok else "FAIL", "graph_ok", graph_ok)
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)        # (NCCL teardown with live graphs hangs; see tensor_p.run_bench)
