"""The full model around the fusion path (SURVEY 8f rank 3 / BASELINE config 5): encoders -> modality dropout -> fusion head ->
classifier and auxiliary heads, with the reference's module tree, `state_dict()` keys, forward signature and output dictionary
(models/multimodal_model.py:12-184), so reference checkpoints load with `strict=True` and `training/advanced_trainer.py` can drive
it unchanged.

What runs where:
  * fusion head, EmotionClassifier, valence / arousal / uncertainty heads, the probability outputs, modality dropout and (when
    `use_sequences=True`) the per-token projection: this package's CUDA library -- no fallback;
  * the three encoders (HF backbone + adapter / prompt / temporal attention / BiLSTM / facial attention / pooling / projection,
    models/encoders.py:10-251): stock PyTorch + Hugging Face modules, as SURVEY 8f says ("encoders stay stock PyTorch/HF").  They
    are outside the hot path; they are restated here only because the reference's own encoder classes call `from_pretrained`
    in their constructors (encoders.py:20,116,179), which needs a network.  `build_backbones()` creates the same architectures
    from their configs with random weights (`AutoModel.from_config`, `Wav2Vec2Model(config)`, `ViTModel(config)`).

Extension (off by default): `use_sequences=True` routes the encoders' `sequence_output` through `SequenceProjector`
(sequence_features.py) so that `mult` / `hierarchical` fusion attend over real token / frame sequences."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from ._lib import B200FusionError
from . import fusion_layers as FL
from .prediction_heads import AuxiliaryHeads, EmotionClassifier
from .sequence_features import SequenceProjector, text_pooling_of

Tensor = torch.Tensor

FUSION_CLASSES = {"early": FL.EarlyFusion, "late": FL.LateFusion, "mult": FL.MultimodalTransformer, "graph": FL.GraphFusion,
                  "contrastive": FL.ContrastiveFusion, "adaptive": FL.AdaptiveFusion, "hierarchical": FL.HierarchicalFusion}
TAKES_CONTRASTIVE_FLAG = ("contrastive", "hierarchical")            # multimodal_model.py:122-135 pass the flag by keyword


# ------------------------------------------------------------------------------------------------------------- backbones
def build_backbones(sizes: str = "base") -> Dict[str, nn.Module]:
    """Random-init backbones of the reference's architectures (config.py:12,17,23: deberta-v3-base, wav2vec2-base, ViT-B/16).
    `sizes='tiny'` keeps the architectures and shrinks every dimension (tests)."""
    from transformers import AutoConfig, AutoModel, ViTConfig, ViTModel, Wav2Vec2Config, Wav2Vec2Model
    if sizes == "base":
        text = AutoConfig.for_model("deberta-v2", vocab_size=128100, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                                    intermediate_size=3072, max_position_embeddings=512, relative_attention=True, position_buckets=256,
                                    norm_rel_ebd="layer_norm", share_att_key=True, pos_att_type=["p2c", "c2p"], position_biased_input=False,
                                    type_vocab_size=0, layer_norm_eps=1e-7)
        audio, video = Wav2Vec2Config(), ViTConfig()
    elif sizes == "tiny":
        text = AutoConfig.for_model("deberta-v2", vocab_size=64, hidden_size=32, num_hidden_layers=1, num_attention_heads=4, intermediate_size=48,
                                    max_position_embeddings=32, relative_attention=True, position_buckets=8, norm_rel_ebd="layer_norm",
                                    share_att_key=True, pos_att_type=["p2c", "c2p"], position_biased_input=False, type_vocab_size=0)
        audio = Wav2Vec2Config(hidden_size=32, num_hidden_layers=1, num_attention_heads=4, intermediate_size=48, conv_dim=(16, 16), conv_stride=(5, 4),
                               conv_kernel=(10, 4), num_feat_extract_layers=2, num_conv_pos_embeddings=8, num_conv_pos_embedding_groups=2,
                               vocab_size=8, mask_time_prob=0.0, mask_feature_prob=0.0)
        video = ViTConfig(hidden_size=32, num_hidden_layers=1, num_attention_heads=4, intermediate_size=48, image_size=16, patch_size=8)
    else:
        raise ValueError("sizes must be 'base' or 'tiny'")
    for c in (text, audio, video):                       # deterministic eval, and training without the backbones' own stochastic masking
        for k in ("hidden_dropout_prob", "attention_probs_dropout_prob", "hidden_dropout", "attention_dropout", "activation_dropout",
                  "feat_proj_dropout", "layerdrop", "final_dropout"):
            if hasattr(c, k):
                setattr(c, k, 0.0)
    return {"text": AutoModel.from_config(text), "audio": Wav2Vec2Model(audio), "video": ViTModel(video, add_pooling_layer=True)}


# ------------------------------------------------------------------------------------------------------------- encoders (stock torch)
class Adapter(nn.Module):
    """bottleneck adapter with a residual connection (encoders.py:254-277)"""

    def __init__(self, hidden: int, bottleneck: int):
        super().__init__()
        self.down_project, self.up_project = nn.Linear(hidden, bottleneck), nn.Linear(bottleneck, hidden)
        self.activation, self.dropout = nn.ReLU(), nn.Dropout(0.1)
        for lin in (self.down_project, self.up_project):
            nn.init.normal_(lin.weight, std=0.02)
            nn.init.zeros_(lin.bias)

    def forward(self, x):
        return x + self.up_project(self.dropout(self.activation(self.down_project(x))))


class _Encoder(nn.Module):
    """What the three encoders share: the optional adapter step and the output dictionary (pooled -> projection -> dropout)."""

    def _adapt(self, seq, use_adapter):
        return self.adapter(seq) if use_adapter and self.adapter is not None else seq

    def _emit(self, pooled, seq, **extra):
        return {"features": self.dropout(self.projection(pooled)), "sequence_output": seq, **extra}


class TextEncoder(_Encoder):
    """encoders.py:10-104 with the backbone passed in; parameter names: model.*, adapter.*, prompt_embeddings, projection.*"""

    def __init__(self, config, backbone: nn.Module):
        super().__init__()
        self.config, self.model = config, backbone
        hidden = backbone.config.hidden_size
        self.hidden_size = hidden
        self.adapter = Adapter(hidden, config.adapter_size) if hasattr(config, "adapter_size") else None
        self.prompt_embeddings = nn.Parameter(torch.randn(config.prompt_length, hidden)) if hasattr(config, "prompt_length") else None
        self.projection = nn.Linear(hidden, config.fusion_hidden_size)
        self.dropout = nn.Dropout(config.fusion_dropout)
        self.pooling = text_pooling_of(backbone.config)

    def forward(self, input_ids, attention_mask, use_adapter: bool = False, use_prompt: bool = False):
        if use_prompt and self.prompt_embeddings is not None:
            B = input_ids.size(0)
            words = self.model.embeddings.word_embeddings(input_ids)
            embeds = torch.cat([self.prompt_embeddings.unsqueeze(0).expand(B, -1, -1), words], dim=1)
            attention_mask = torch.cat([attention_mask.new_ones(B, self.prompt_embeddings.size(0)), attention_mask], dim=1)
            seq = self.model(inputs_embeds=embeds, attention_mask=attention_mask).last_hidden_state
        else:
            seq = self.model(input_ids=input_ids, attention_mask=attention_mask).last_hidden_state
        seq = self._adapt(seq, use_adapter)
        if self.pooling == "cls":
            pooled = seq[:, 0]
        else:
            m = attention_mask.unsqueeze(-1).expand(seq.size())
            pooled = (seq * m).sum(1) / m.sum(1).clamp(min=1e-9)
        return self._emit(pooled, seq, attention_mask=attention_mask)


class AudioEncoder(_Encoder):
    """encoders.py:107-167: wav2vec2 -> adapter -> temporal self-attention -> mean over time -> projection"""

    def __init__(self, config, backbone: nn.Module):
        super().__init__()
        self.config, self.model = config, backbone
        hidden = backbone.config.hidden_size
        self.hidden_size = hidden
        self.adapter = Adapter(hidden, config.adapter_size) if hasattr(config, "adapter_size") else None
        self.temporal_attention = nn.MultiheadAttention(hidden, num_heads=8, dropout=config.fusion_dropout, batch_first=True)
        self.projection = nn.Linear(hidden, config.fusion_hidden_size)
        self.dropout = nn.Dropout(config.fusion_dropout)

    def forward(self, waveform, use_adapter: bool = False):
        seq = self._adapt(self.model(waveform).last_hidden_state, use_adapter)
        attended, weights = self.temporal_attention(seq, seq, seq)
        return self._emit(attended.mean(dim=1), attended, attention_weights=weights)


class VideoEncoder(_Encoder):
    """encoders.py:170-251: ViT CLS per frame -> adapter -> 2-layer BiLSTM -> facial self-attention -> mean over frames -> projection"""

    def __init__(self, config, backbone: nn.Module):
        super().__init__()
        self.config, self.vit = config, backbone
        hidden = backbone.config.hidden_size
        self.hidden_size = hidden
        self.temporal_lstm = nn.LSTM(hidden, hidden // 2, num_layers=2, batch_first=True, bidirectional=True, dropout=config.fusion_dropout)
        self.facial_attention = nn.MultiheadAttention(hidden, num_heads=8, dropout=config.fusion_dropout, batch_first=True)
        self.adapter = Adapter(hidden, config.adapter_size) if hasattr(config, "adapter_size") else None
        self.projection = nn.Linear(hidden, config.fusion_hidden_size)
        self.dropout = nn.Dropout(config.fusion_dropout)

    def forward(self, video_frames, use_adapter: bool = False):
        B, F_, C, Hh, Ww = video_frames.shape
        cls = self.vit(pixel_values=video_frames.reshape(B * F_, C, Hh, Ww)).last_hidden_state[:, 0]
        frames = self._adapt(cls.reshape(B, F_, -1), use_adapter)
        rec, _ = self.temporal_lstm(frames)
        attended, weights = self.facial_attention(rec, rec, rec)
        return self._emit(attended.mean(dim=1), attended, attention_weights=weights)


# ------------------------------------------------------------------------------------------------------------- the model
class MultimodalEmotionModel(nn.Module):
    """MultimodalEmotionModel(config, backbones=None, use_sequences=False) -- reference models/multimodal_model.py:12-184.

    `backbones`: {'text', 'audio', 'video'} -> nn.Module (default: `build_backbones('base')`, random init)."""

    def __init__(self, config, backbones: Optional[Dict[str, nn.Module]] = None, use_sequences: bool = False):
        super().__init__()
        self.config = config
        bb = backbones if backbones is not None else build_backbones("base")
        self.text_encoder = TextEncoder(config, bb["text"])
        self.audio_encoder = AudioEncoder(config, bb["audio"])
        self.video_encoder = VideoEncoder(config, bb["video"])
        self.modality_dropout = FL.ModalityDropout(dropout_rate=0.1)
        self.fusion_type = getattr(config, "fusion_type", "hierarchical")
        if self.fusion_type not in FUSION_CLASSES:
            raise ValueError(f"Unknown fusion type: {self.fusion_type}")
        self.fusion_layer = FUSION_CLASSES[self.fusion_type](config)
        self.classifier = None if self.fusion_type == "late" else EmotionClassifier(config)
        aux = AuxiliaryHeads(config)
        # the three auxiliary Linear layers live at the top level of the reference's module tree (multimodal_model.py:55-60);
        # `aux` evaluates them as one GEMM and is kept outside the module registry so each parameter has exactly one name
        self.valence_regressor, self.arousal_regressor, self.uncertainty_head = aux.valence_regressor, aux.arousal_regressor, aux.uncertainty_head
        object.__setattr__(self, "_aux", aux)
        self.use_sequences = bool(use_sequences)
        if self.use_sequences:
            if self.fusion_type not in ("mult", "hierarchical"):
                raise B200FusionError("use_sequences=True needs a fusion head that attends over sequences ('mult' or 'hierarchical')")
            sp = SequenceProjector(config, self.text_encoder.projection, self.audio_encoder.projection, self.video_encoder.projection,
                                   text_pooling=self.text_encoder.pooling)
            object.__setattr__(self, "_sequence_projector", sp)       # shares the encoders' projection parameters: nothing to register

    def train(self, mode: bool = True):
        super().train(mode)
        self._aux.train(mode)
        if self.use_sequences:
            self._sequence_projector.train(mode)
        return self

    @staticmethod
    def _blank(text_input, audio_input, video_input, missing: Optional[Sequence[str]]):
        """zero the inputs of modalities listed as missing (multimodal_model.py:77-87)"""
        if missing:
            if "text" in missing:
                text_input = {k: torch.zeros_like(text_input[k]) for k in ("input_ids", "attention_mask")}
            if "audio" in missing:
                audio_input = torch.zeros_like(audio_input)
            if "video" in missing:
                video_input = torch.zeros_like(video_input)
        return text_input, audio_input, video_input

    def forward(self, text_input: Dict[str, Tensor], audio_input: Tensor, video_input: Tensor, use_adapter: bool = False, use_prompt: bool = False,
                compute_contrastive_loss: bool = False, missing_modalities: Optional[List[str]] = None) -> Dict[str, Tensor]:
        text_input, audio_input, video_input = self._blank(text_input, audio_input, video_input, missing_modalities)
        enc_t = self.text_encoder(text_input["input_ids"], text_input["attention_mask"], use_adapter=use_adapter, use_prompt=use_prompt)
        enc_a = self.audio_encoder(audio_input, use_adapter=use_adapter)
        enc_v = self.video_encoder(video_input, use_adapter=use_adapter)
        feats = (enc_t["features"], enc_a["features"], enc_v["features"])
        fusion_kw = {"compute_contrastive_loss": compute_contrastive_loss} if self.fusion_type in TAKES_CONTRASTIVE_FLAG else {}
        if self.use_sequences:
            seq = self._sequence_projector(enc_t, enc_a, enc_v)
            feats = (seq["text_features"], seq["audio_features"], seq["video_features"])
            mask = self.modality_dropout.sample_mask(feats[0].size(0), feats[0].device) if self.training else None
            feats = FL._masked_split(*feats, mask)                      # the pooled features the model returns carry the keep-mask too
            fusion_in = (seq["text_sequence"], seq["audio_sequence"], seq["video_sequence"])
            if self.fusion_type == "hierarchical":
                fusion_kw["pooled_features"] = feats
            fusion_output = self.fusion_layer(*fusion_in, mask=mask, **fusion_kw)
        else:
            if self.training:
                feats = self.modality_dropout(*feats, training=True)
            fusion_output = self.fusion_layer(*feats, **fusion_kw)
        t, a, v = feats

        if self.fusion_type == "late":
            emotion_logits = fusion_output["fused_logits"]
            head_in = (t + a + v) / 3                                   # late fusion has no fused vector (multimodal_model.py:152-157)
        else:
            head_in = fusion_output["fused_features"] if isinstance(fusion_output, dict) else fusion_output
            emotion_logits = self.classifier(head_in)
        aux = self._aux(head_in, emotion_logits)
        output = {"emotion_logits": emotion_logits, "emotion_probs": aux["emotion_probs"], "valence": aux["valence"], "arousal": aux["arousal"],
                  "uncertainty": aux["uncertainty"], "text_features": t, "audio_features": a, "video_features": v}
        if self.fusion_type == "late":
            output["individual_logits"] = {m: fusion_output[f"{m}_logits"] for m in ("text", "audio", "video")}
            output["fusion_weights"] = fusion_output["fusion_weights"]
        if isinstance(fusion_output, dict):                            # every other fusion output is passed through, and may overwrite the
            output.update({k: x for k, x in fusion_output.items() if k != "fused_features"})   # encoder features (mult; :177-181)
        return output


def load_pretrained_model(checkpoint_path: str, config, backbones=None) -> MultimodalEmotionModel:
    """multimodal_model.py:472-485: accepts a bare state_dict or a trainer checkpoint with 'model_state_dict'"""
    model = MultimodalEmotionModel(config, backbones)
    ckpt = torch.load(checkpoint_path, map_location="cpu")
    model.load_state_dict(ckpt["model_state_dict"] if "model_state_dict" in ckpt else ckpt)
    return model
