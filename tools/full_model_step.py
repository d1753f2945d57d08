"""BASELINE config 5 (not a pytest, not bench.py's metric), on one GPU or data-parallel under torchrun: the full model -- random-init DeBERTa-v3-base, Wav2Vec2-base and
ViT-B/16 + BiLSTM encoders (stock PyTorch/HF, bf16 autocast) feeding hierarchical fusion and the heads (this package's CUDA library) --
one training step = forward, the trainer's loss (label-smoothed CE + 0.1 * sum InfoNCE, training/advanced_trainer.py:139-166),
backward, clip_grad_norm_(1.0) + AdamW (FusedAdamW).  Times the step with CUDA events and, separately, the fusion + heads part of it
(every C-ABI call bracketed by events), so the share of the step spent on the fusion path is measured, not guessed.

    python tools/full_model_step.py [--batch 8] [--steps 5] [--sequences]  > gpurun_out/full_model_step.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/full_model_step.py --batch 8

Data-parallel (BASELINE configs[4]: "8 B200 data-parallel"): one process per GPU, per-GPU batch fixed, every parameter gradient in one
flat fp32 `GradBucket` all-reduced in place once per step (SUM; the local-mean cross-entropy is scaled by 1/world, the InfoNCE terms
already span the global batch through the all-gather inside the head), then clip + AdamW on every rank.  Step time = max over ranks.
"""
import argparse
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simple-multimodal_b200")
em = importlib.import_module("simple-multimodal_b200.emotion_model")

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--text-len", type=int, default=512)
ap.add_argument("--audio-samples", type=int, default=160000)
ap.add_argument("--frames", type=int, default=30)
ap.add_argument("--sequences", action="store_true", help="use_sequences=True: MulT attends over the projected token / frame sequences")
args = ap.parse_args()


class Cfg:                                         # reference config.py defaults, graph_hidden_size = fusion_hidden_size
    fusion_hidden_size, fusion_dropout, fusion_num_heads, num_emotions = 512, 0.1, 8, 7
    graph_hidden_size, graph_num_layers, graph_dropout, contrastive_temperature = 512, 3, 0.1, 0.07
    adapter_size, prompt_length, fusion_type = 64, 10, "hierarchical"


import torch.distributed as dist                                                   # noqa: E402

world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)                                                               # identical random-init weights on every rank
t0 = time.perf_counter()
model = em.MultimodalEmotionModel(Cfg, em.build_backbones("base"), use_sequences=args.sequences).to(dev).train()
build_s = time.perf_counter() - t0
n_params = sum(p.numel() for p in model.parameters())
n_fusion = sum(p.numel() for n, p in model.named_parameters() if n.startswith(("fusion_layer.", "classifier.", "valence", "arousal", "uncertainty")))
opt = pkg.FusedAdamW([{"params": [p for n, p in model.named_parameters() if ".model." in n or ".vit." in n], "lr": 1e-5},
                      {"params": [p for n, p in model.named_parameters() if not (".model." in n or ".vit." in n)], "lr": 1e-4}], weight_decay=1e-5)
ce = pkg.SmoothedCrossEntropy(0.1)
bucket = pkg.GradBucket(model.parameters()) if world > 1 else None
g = torch.Generator().manual_seed(1 + rank)                                        # every rank its own batch
B = args.batch
host = {"ids": torch.randint(0, 128100, (B, args.text_len), generator=g).pin_memory(), "am": torch.ones(B, args.text_len, dtype=torch.long).pin_memory(),
        "audio": torch.randn(B, args.audio_samples, generator=g).pin_memory(), "video": torch.randn(B, args.frames, 3, 224, 224, generator=g).pin_memory(),
        "labels": torch.randint(0, 7, (B,), generator=g).pin_memory()}


def step():
    x = {k: v.to(dev, non_blocking=True) for k, v in host.items()}                  # H2D of the raw inputs every step
    if bucket is None:
        opt.zero_grad(set_to_none=True)
    else:
        bucket.zero()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model({"input_ids": x["ids"], "attention_mask": x["am"]}, x["audio"], x["video"], compute_contrastive_loss=True)
    loss = ce(out["emotion_logits"], x["labels"]) * pkg.global_batch_scale() + 0.1 * sum(out["contrastive_losses"].values())
    loss.backward()
    if bucket is not None:
        bucket.all_reduce()
    opt.clip_grad_norm_(1.0)
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
if world > 1:
    tt = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt[0])
# the library's share: bracket every C-ABI call of one more step
pkg._lib.CALL_PROFILE = []
step()
torch.cuda.synchronize()
prof, pkg._lib.CALL_PROFILE = pkg._lib.CALL_PROFILE, None
lib_ms = sum(a.elapsed_time(b) for _, a, b in prof)
if world > 1:
    dist.barrier()
if rank == 0:
  print(json.dumps({"workload": "config 5: full model, random-init deberta-v3-base + wav2vec2-base + ViT-B/16, hierarchical fusion",
                  "n_gpus": world, "parallelism": f"dp{world}", "global_batch": B * world, "samples_per_s_all_gpus": B * world / (ms * 1e-3),
                  "gradient_exchange": "one in-place all-reduce of a flat fp32 GradBucket" if world > 1 else "none",
                  "use_sequences": args.sequences, "batch": B, "text_len": args.text_len, "audio_samples": args.audio_samples, "frames": args.frames,
                  "params_total": n_params, "params_fusion_and_heads": n_fusion, "model_build_s": build_s, "ms_per_step": ms,
                  "samples_per_s": B / (ms * 1e-3), "library_calls_per_step": len(prof), "library_ms_per_step": lib_ms,
                  "library_share_of_step": lib_ms / ms, "loss": float(loss.detach()), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                  "encoders": "stock PyTorch/HF under bf16 autocast", "optimizer": "FusedAdamW (clip 1.0), two param groups"}))
if world > 1:
    dist.destroy_process_group()
