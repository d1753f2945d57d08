"""Host-side logic that needs no GPU: the optimizer's device chunk table (layout documented in include/b200_fusion.h), the
pooling rule the text encoder applies per backbone type, refusal of CPU tensors by the extension modules (no fallback)."""
import importlib

import pytest
import torch
import torch.nn as nn

pkg = importlib.import_module("simple-multimodal_b200")
optim = importlib.import_module("simple-multimodal_b200.optim")


def test_chunk_table_layout():
    shapes = [(5000,), (3, 7), (4096,), (1,)]
    ps = [torch.zeros(s) for s in shapes]
    gs, ms, vs = ([torch.zeros(s) for s in shapes] for _ in range(3))
    tab = optim._Table(ps, gs, ms, vs, "cpu")
    T = len(shapes)
    assert tab.n_tensors == T and tab.n_chunks == 2 + 1 + 1 + 1
    raw = tab.buf.numpy().tobytes()
    head = torch.frombuffer(bytearray(raw[:T * 40]), dtype=torch.int64)
    assert head[:T].tolist() == [t.data_ptr() for t in ps]
    assert head[T:2 * T].tolist() == [t.data_ptr() for t in gs]
    assert head[2 * T:3 * T].tolist() == [t.data_ptr() for t in ms]
    assert head[3 * T:4 * T].tolist() == [t.data_ptr() for t in vs]
    assert head[4 * T:5 * T].tolist() == [5000, 21, 4096, 1]
    chunks = torch.frombuffer(bytearray(raw[T * 40:]), dtype=torch.int32).view(-1, 2).tolist()
    assert chunks == [[0, 0], [0, 4096], [1, 0], [2, 0], [3, 0]]          # (tensor, first element), 4096 elements per chunk


def test_fused_adamw_keeps_torch_optimizer_semantics_on_the_host():
    p = [nn.Parameter(torch.zeros(4)), nn.Parameter(torch.zeros(2, 2))]
    opt = pkg.FusedAdamW([{"params": p[:1], "lr": 1e-4}, {"params": p[1:]}], lr=1e-3, weight_decay=0.01)
    assert [g["lr"] for g in opt.param_groups] == [1e-4, 1e-3] and all(g["weight_decay"] == 0.01 for g in opt.param_groups)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=[1e-4, 1e-3], total_steps=10)     # schedulers only touch param_groups
    assert len(sched.get_last_lr()) == 2
    opt.step()                                                             # no gradients anywhere: nothing to launch, nothing raised
    assert float(opt.clip_grad_norm_(1.0)) == 0.0


@pytest.mark.parametrize("model_type,want", [("deberta-v2", "cls"), ("bert", "cls"), ("roberta", "cls"), ("electra", "masked_mean"), (None, "masked_mean")])
def test_text_pooling_rule_follows_the_reference(model_type, want):
    class C:
        pass
    c = C()
    if model_type is not None:
        c.model_type = model_type
    assert pkg.text_pooling_of(c) == want                                   # reference models/encoders.py:86: 'bert' in model_type


def test_sequence_projector_refuses_cpu_tensors():
    class Cfg:
        fusion_hidden_size, fusion_dropout = 16, 0.0
    sp = pkg.SequenceProjector(Cfg, nn.Linear(8, 16), nn.Linear(8, 16), nn.Linear(8, 16), text_pooling="masked_mean")
    enc = {"sequence_output": torch.zeros(2, 3, 8), "attention_mask": torch.ones(2, 3, dtype=torch.long)}
    with pytest.raises(pkg.B200FusionError):
        sp(enc, enc, enc)
    with pytest.raises(ValueError):
        pkg.SequenceProjector(Cfg, nn.Linear(8, 16), nn.Linear(8, 16), nn.Linear(8, 16), text_pooling="max")


def test_sign_bit_buffers_and_stash_accounting():
    """One-bit ReLU' mask (b200f_gemm_args::sign_bits*): the buffer exists only where the bf16 tcgen05 epilogue can write it, and the
    resident-stash estimate of the MulT engine counts it."""
    K = pkg.kernels
    x = torch.zeros(96, 512, dtype=torch.bfloat16)
    bits = K.sign_bits_for(x, 2048)
    assert bits.dtype == torch.int32 and tuple(bits.shape) == (96, 64) and bits.is_contiguous()
    assert K.sign_bits_for(x.float(), 2048) is None                       # fp32 parity mode keeps relu_mask=
    assert K.sign_bits_for(x, 2000) is None and K.sign_bits_for(x[:, :100], 2048) is None
    me = importlib.import_module("simple-multimodal_b200.mult_engine")
    Ls, H = (512, 512, 30), 512
    bf16, fp32 = me.stash_bytes_per_sample(Ls, H, 2), me.stash_bytes_per_sample(Ls, H, 4)
    assert bf16 - fp32 // 2 == 2 * (4 * H // 8) * sum(Ls)                  # two hidden layers per token, one bit per element


def test_bind_host_to_gpu_is_a_hint_only():
    """No NVML device behind a CPU 'device': the affinity helper must leave the process alone and not raise."""
    import os
    before = os.sched_getaffinity(0)
    assert pkg.bind_host_to_gpu("cpu") == []
    assert os.sched_getaffinity(0) == before
