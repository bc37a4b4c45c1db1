"""Plain-PyTorch fp32 restatement of the reference's encoder-side hot path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: SpeechBrain
is an un-vendored dependency of the reference and cannot be imported here, so
every class below restates the published SpeechBrain ~v0.5.14 algorithm and is
anchored on the reference's own call sites:

  call sequence            /root/reference/stac-st/inference.py:95-107
                           /root/reference/stac-st/train_multitask.py:59-78
  hyper-parameters         /root/reference/stac-st/hparams/transformer_multitask.yaml:161-210,253-254,299-302
  glue (encode / masks)    /root/reference/stac-st/modules/TransformerMultiTask.py:90-142,211-232,273-314

The module tree and ``state_dict`` key layout mirror SpeechBrain's (SURVEY.md
Appendix A.7) so that a reference checkpoint would load unchanged.  The cost
structure is deliberately the reference's too (six eager stages, ``torch.stft``,
``F.pad(reflect)`` + ``nn.Conv2d``, ``nn.MultiheadAttention`` slow path with
``need_weights=True``) because this file doubles as the timed CPU baseline.

Assumptions that only SpeechBrain's source pins (kept switchable where cheap):
  * top-dB clamp is per utterance (``Filterbank._amplitude_to_DB`` uses
    ``amax(dim=(-2,-1))``); ``top_db_per_utterance=False`` gives the older
    batch-global behaviour.
  * ``InputNormalization``: mean_norm=True, std_norm=True, unbiased std, eps 1e-10.
  * ``Conv2d``: padding="same" with stride>1 -> pad k//2 both sides, reflect mode.
  * CNN LayerNorm normalises over (freq, channel), eps 1e-5; Transformer LayerNorm eps 1e-6.
  * exact-erf GELU, LeakyReLU slope 0.01, no sqrt(d) scaling on the source side.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

# (d_model, nhead, num_encoder_layers, d_ffn)
# /root/reference/run_default.sh:73-76 ; /root/reference/ablations/run_m_and_l_size.sh:72-99
MODEL_SIZES = {
    "S": dict(d_model=256, nhead=4, num_encoder_layers=12, d_ffn=1024),
    "M": dict(d_model=512, nhead=8, num_encoder_layers=16, d_ffn=2048),
    "L": dict(d_model=1024, nhead=16, num_encoder_layers=14, d_ffn=4096),
}


# --------------------------------------------------------------------------
# speechbrain.processing.features.{STFT, spectral_magnitude, Filterbank}
# speechbrain.lobes.features.Fbank
# --------------------------------------------------------------------------
class STFT(nn.Module):
    """speechbrain/processing/features.py::STFT (win 25 ms, hop 10 ms, hamming)."""

    def __init__(self, sample_rate, win_length=25, hop_length=10, n_fft=400,
                 window_fn=torch.hamming_window, normalized_stft=False,
                 center=True, pad_mode="constant", onesided=True):
        super().__init__()
        self.sample_rate = sample_rate
        self.n_fft = n_fft
        self.normalized_stft = normalized_stft
        self.center = center
        self.pad_mode = pad_mode
        self.onesided = onesided
        self.win_length = int(round((sample_rate / 1000.0) * win_length))
        self.hop_length = int(round((sample_rate / 1000.0) * hop_length))
        self.window = window_fn(self.win_length)

    def forward(self, x):
        stft = torch.stft(
            x, self.n_fft, self.hop_length, self.win_length,
            self.window.to(x.device), self.center, self.pad_mode,
            self.normalized_stft, self.onesided, return_complex=True,
        )
        stft = torch.view_as_real(stft)
        return stft.transpose(2, 1)  # [B, T, n_fft//2+1, 2]


def spectral_magnitude(stft, power=1, log=False, eps=1e-14):
    """speechbrain/processing/features.py::spectral_magnitude (power=1 -> re^2+im^2)."""
    spectr = stft.pow(2).sum(-1)
    if power < 1:
        spectr = spectr + eps
    spectr = spectr.pow(power)
    if log:
        return torch.log(spectr + eps)
    return spectr


class Filterbank(nn.Module):
    """speechbrain/processing/features.py::Filterbank (triangular, log-mel, top_db 80).

    The triangle of filter i is symmetric in Hz with half-width equal to the
    LEFT mel-band spacing ``hz[i+1]-hz[i]`` (SpeechBrain's variant; not HTK/Slaney).
    """

    def __init__(self, n_mels=40, log_mel=True, f_min=0, f_max=8000, n_fft=400,
                 sample_rate=16000, power_spectrogram=2, amin=1e-10,
                 ref_value=1.0, top_db=80.0, top_db_per_utterance=True):
        super().__init__()
        self.n_mels = n_mels
        self.log_mel = log_mel
        self.f_min = f_min
        self.f_max = f_max
        self.n_fft = n_fft
        self.sample_rate = sample_rate
        self.amin = amin
        self.ref_value = ref_value
        self.top_db = top_db
        self.top_db_per_utterance = top_db_per_utterance
        self.n_stft = n_fft // 2 + 1
        self.db_multiplier = math.log10(max(amin, ref_value))
        self.multiplier = 10 if power_spectrogram == 2 else 20

        mel = torch.linspace(self._to_mel(f_min), self._to_mel(f_max), n_mels + 2)
        hz = self._to_hz(mel)
        band = hz[1:] - hz[:-1]
        self.band = band[:-1]
        self.f_central = hz[1:-1]
        all_freqs = torch.linspace(0, sample_rate // 2, self.n_stft)
        self.all_freqs_mat = all_freqs.repeat(self.f_central.shape[0], 1)

    @staticmethod
    def _to_mel(hz):
        return 2595 * math.log10(1 + hz / 700)

    @staticmethod
    def _to_hz(mel):
        return 700 * (10 ** (mel / 2595) - 1)

    def fbank_matrix(self):
        """[n_stft, n_mels] filter matrix; SpeechBrain rebuilds it on every call."""
        f_central_mat = self.f_central.repeat(self.all_freqs_mat.shape[1], 1).transpose(0, 1)
        band_mat = self.band.repeat(self.all_freqs_mat.shape[1], 1).transpose(0, 1)
        slope = (self.all_freqs_mat - f_central_mat) / band_mat
        left_side = slope + 1.0
        right_side = -slope + 1.0
        zero = torch.zeros(1)
        return torch.max(zero, torch.min(left_side, right_side)).transpose(0, 1)

    def forward(self, spectrogram):
        fbank_matrix = self.fbank_matrix().to(spectrogram.device)
        fbanks = torch.matmul(spectrogram, fbank_matrix)
        if self.log_mel:
            fbanks = self._amplitude_to_DB(fbanks)
        return fbanks

    def _amplitude_to_DB(self, x):
        x_db = self.multiplier * torch.log10(torch.clamp(x, min=self.amin))
        x_db = x_db - self.multiplier * self.db_multiplier
        if self.top_db_per_utterance:
            new_x_db_max = x_db.amax(dim=(-2, -1)) - self.top_db
            x_db = torch.max(x_db, new_x_db_max.view(x_db.shape[0], 1, 1))
        else:
            x_db = torch.max(x_db, x_db.max() - self.top_db)
        return x_db


class Fbank(nn.Module):
    """speechbrain/lobes/features.py::Fbank as configured at
    /root/reference/stac-st/hparams/transformer_multitask.yaml:299-302
    (sample_rate=16000, n_fft=400, n_mels=80; deltas/context off).
    """

    def __init__(self, sample_rate=16000, f_min=0, f_max=None, n_fft=400,
                 n_mels=40, win_length=25, hop_length=10,
                 top_db_per_utterance=True):
        super().__init__()
        if f_max is None:
            f_max = sample_rate / 2
        self.compute_STFT = STFT(sample_rate=sample_rate, n_fft=n_fft,
                                 win_length=win_length, hop_length=hop_length)
        self.compute_fbanks = Filterbank(sample_rate=sample_rate, n_fft=n_fft,
                                         n_mels=n_mels, f_min=f_min, f_max=f_max,
                                         top_db_per_utterance=top_db_per_utterance)

    def forward(self, wav):
        stft = self.compute_STFT(wav)
        mag = spectral_magnitude(stft)
        return self.compute_fbanks(mag)


# --------------------------------------------------------------------------
# speechbrain.processing.features.InputNormalization
# --------------------------------------------------------------------------
class InputNormalization(nn.Module):
    """speechbrain/processing/features.py::InputNormalization (norm_type="global").

    yaml: /root/reference/stac-st/hparams/transformer_multitask.yaml:208-210.
    Statistics are plain attributes (not buffers), saved by SpeechBrain's
    checkpointer through ``_statistics_dict`` (normalizer.ckpt).
    """

    def __init__(self, mean_norm=True, std_norm=True, norm_type="global",
                 avg_factor=None, update_until_epoch=3):
        super().__init__()
        if norm_type != "global":
            raise NotImplementedError("the reference path only uses norm_type=global")
        self.mean_norm = mean_norm
        self.std_norm = std_norm
        self.norm_type = norm_type
        self.avg_factor = avg_factor
        self.update_until_epoch = update_until_epoch
        self.glob_mean = torch.tensor([0])
        self.glob_std = torch.tensor([0])
        self.weight = 1.0
        self.count = 0
        self.eps = 1e-10

    def _compute_current_stats(self, x):
        if self.mean_norm:
            current_mean = torch.mean(x, dim=0).detach()
        else:
            current_mean = torch.tensor([0.0], device=x.device)
        if self.std_norm:
            current_std = torch.std(x, dim=0).detach()
        else:
            current_std = torch.tensor([1.0], device=x.device)
        current_std = torch.max(current_std, self.eps * torch.ones_like(current_std))
        return current_mean, current_std

    def forward(self, x, lengths, spk_ids=torch.tensor([]), epoch=0):
        n_batches = x.shape[0]
        current_means, current_stds = [], []
        for snt_id in range(n_batches):  # Python loop, also in eval (as in SpeechBrain)
            actual_size = torch.round(lengths[snt_id] * x.shape[1]).int()
            m, s = self._compute_current_stats(x[snt_id, 0:actual_size, ...])
            current_means.append(m)
            current_stds.append(s)
        current_mean = torch.mean(torch.stack(current_means), dim=0)
        current_std = torch.mean(torch.stack(current_stds), dim=0)
        if self.training:
            if self.count == 0:
                self.glob_mean = current_mean
                self.glob_std = current_std
            elif epoch < self.update_until_epoch:
                self.weight = 1 / (self.count + 1) if self.avg_factor is None else self.avg_factor
                self.glob_mean = (1 - self.weight) * self.glob_mean + self.weight * current_mean
                self.glob_std = (1 - self.weight) * self.glob_std + self.weight * current_std
            self.count = self.count + 1
        return (x - self.glob_mean.data) / self.glob_std.data

    def _statistics_dict(self):
        return {"count": self.count, "glob_mean": self.glob_mean, "glob_std": self.glob_std,
                "spk_dict_mean": {}, "spk_dict_std": {}, "spk_dict_count": {}}

    def _load_statistics_dict(self, state):
        self.count = state["count"]
        self.glob_mean = state["glob_mean"]
        self.glob_std = state["glob_std"]
        return state


# --------------------------------------------------------------------------
# speechbrain.nnet.{CNN.Conv2d, normalization.LayerNorm, linear.Linear}
# speechbrain.lobes.models.convolution.{ConvBlock, ConvolutionFrontEnd}
# --------------------------------------------------------------------------
class Conv2d(nn.Module):
    """speechbrain/nnet/CNN.py::Conv2d, padding="same", padding_mode="reflect".

    Input [B,T,F] (-> one channel) or [B,T,F,C]; internally [B,C,F,T] so H=freq,
    W=time; "same" with stride>1 pads kernel//2 on both sides of both axes.
    """

    def __init__(self, out_channels, kernel_size, in_channels, stride=(1, 1),
                 unsqueeze=False, bias=True):
        super().__init__()
        if isinstance(kernel_size, int):
            kernel_size = (kernel_size, kernel_size)
        if isinstance(stride, int):
            stride = (stride, stride)
        self.kernel_size = kernel_size
        self.stride = stride
        self.unsqueeze = unsqueeze
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                              padding=0, bias=bias)

    def forward(self, x):
        x = x.transpose(1, -1)
        if self.unsqueeze:
            x = x.unsqueeze(1)
        pads = []
        for k, s in ((self.kernel_size[-1], self.stride[-1]), (self.kernel_size[-2], self.stride[-2])):
            if s <= 1:
                raise NotImplementedError("reference path uses stride 2 only")
            pads += [k // 2, k // 2]  # get_padding_elem, stride > 1 branch
        x = nn.functional.pad(x, pads, mode="reflect")
        wx = self.conv(x)
        if self.unsqueeze:
            wx = wx.squeeze(1)
        return wx.transpose(1, -1)


class LayerNorm(nn.Module):
    """speechbrain/nnet/normalization.py::LayerNorm (wraps torch LayerNorm as ``.norm``)."""

    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
        super().__init__()
        self.norm = nn.LayerNorm(normalized_shape, eps=eps, elementwise_affine=elementwise_affine)

    def forward(self, x):
        return self.norm(x)


class Linear(nn.Module):
    """speechbrain/nnet/linear.py::Linear (wraps nn.Linear as ``.w``)."""

    def __init__(self, n_neurons, input_size, bias=True, combine_dims=False):
        super().__init__()
        self.combine_dims = combine_dims
        self.w = nn.Linear(input_size, n_neurons, bias=bias)

    def forward(self, x):
        if x.ndim == 4 and self.combine_dims:
            x = x.reshape(x.shape[0], x.shape[1], x.shape[2] * x.shape[3])
        return self.w(x)


class _Named(nn.Module):
    """Stand-in for speechbrain.nnet.containers.Sequential: named children run in order."""

    def __init__(self, **layers):
        super().__init__()
        for name, layer in layers.items():
            self.add_module(name, layer)

    def forward(self, x):
        for layer in self.children():
            x = layer(x)
        return x


class ConvBlock(nn.Module):
    """speechbrain/lobes/models/convolution.py::ConvBlock, one layer, no residual."""

    def __init__(self, in_channels, in_freq, out_channels, kernel_size, stride, dropout, unsqueeze):
        super().__init__()
        out_freq = (in_freq - 1) // stride + 1
        self.convs = _Named(
            conv_0=Conv2d(out_channels, kernel_size, in_channels, stride=stride, unsqueeze=unsqueeze),
            norm_0=LayerNorm((out_freq, out_channels), eps=1e-5),
            act_0=nn.LeakyReLU(),
        )
        self.drop = nn.Dropout(dropout)
        self.out_freq = out_freq

    def forward(self, x):
        return self.drop(self.convs(x))


class ConvolutionFrontEnd(nn.Module):
    """speechbrain/lobes/models/convolution.py::ConvolutionFrontEnd as configured at
    /root/reference/stac-st/hparams/transformer_multitask.yaml:173-180.
    [B,T,80] -> [B,T',40,256] -> [B,T'',20,256].
    """

    def __init__(self, input_shape, num_blocks=3, num_layers_per_block=5,
                 out_channels=(128, 256, 512), kernel_sizes=(3, 3, 3),
                 strides=(1, 2, 2), residuals=(True, True, True), dropout=0.1):
        super().__init__()
        if num_layers_per_block != 1 or any(residuals):
            raise NotImplementedError("reference path: 1 layer per block, no residuals")
        freq, chans = input_shape[-1], 1
        for i in range(num_blocks):
            block = ConvBlock(chans, freq, out_channels[i], kernel_sizes[i], strides[i],
                              dropout, unsqueeze=(i == 0))
            self.add_module(f"convblock_{i}", block)
            freq, chans = block.out_freq, out_channels[i]

    def forward(self, x):
        for block in self.children():
            x = block(x)
        return x


# --------------------------------------------------------------------------
# speechbrain.lobes.models.transformer.Transformer.{PositionalEncoding,
#   TransformerEncoderLayer, TransformerEncoder}; speechbrain.nnet.attention.*
# --------------------------------------------------------------------------
class PositionalEncoding(nn.Module):
    def __init__(self, input_size, max_len=2500):
        super().__init__()
        pe = torch.zeros(max_len, input_size)
        positions = torch.arange(0, max_len).unsqueeze(1).float()
        denominator = torch.exp(
            torch.arange(0, input_size, 2).float() * -(math.log(10000.0) / input_size))
        pe[:, 0::2] = torch.sin(positions * denominator)
        pe[:, 1::2] = torch.cos(positions * denominator)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        return self.pe[:, : x.size(1)].clone().detach()


class MultiheadAttention(nn.Module):
    """speechbrain/nnet/attention.py::MultiheadAttention: nn.MultiheadAttention with
    batch_first=False and need_weights=True (torch's slow path; weights head-averaged)."""

    def __init__(self, nhead, d_model, dropout=0.0):
        super().__init__()
        self.att = nn.MultiheadAttention(embed_dim=d_model, num_heads=nhead, dropout=dropout, bias=True)

    def forward(self, query, key, value, attn_mask=None, key_padding_mask=None):
        query, key, value = (t.permute(1, 0, 2) for t in (query, key, value))
        output, attention_weights = self.att(
            query, key, value, attn_mask=attn_mask,
            key_padding_mask=key_padding_mask, need_weights=True)
        return output.permute(1, 0, 2), attention_weights


class PositionalwiseFeedForward(nn.Module):
    def __init__(self, d_ffn, input_size, dropout=0.0, activation=nn.ReLU):
        super().__init__()
        self.ffn = nn.Sequential(
            nn.Linear(input_size, d_ffn), activation(), nn.Dropout(dropout),
            nn.Linear(d_ffn, input_size))

    def forward(self, x):
        x = x.permute(1, 0, 2)
        x = self.ffn(x)
        return x.permute(1, 0, 2)


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_ffn, nhead, d_model, dropout, activation, normalize_before):
        super().__init__()
        self.self_att = MultiheadAttention(nhead=nhead, d_model=d_model, dropout=dropout)
        self.pos_ffn = PositionalwiseFeedForward(d_ffn=d_ffn, input_size=d_model,
                                                 dropout=dropout, activation=activation)
        self.norm1 = LayerNorm(d_model, eps=1e-6)
        self.norm2 = LayerNorm(d_model, eps=1e-6)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.normalize_before = normalize_before

    def forward(self, src, src_mask=None, src_key_padding_mask=None, pos_embs=None):
        src1 = self.norm1(src) if self.normalize_before else src
        output, self_attn = self.self_att(src1, src1, src1, attn_mask=src_mask,
                                          key_padding_mask=src_key_padding_mask)
        src = src + self.dropout1(output)
        if not self.normalize_before:
            src = self.norm1(src)
        src1 = self.norm2(src) if self.normalize_before else src
        output = self.pos_ffn(src1)
        output = src + self.dropout2(output)
        if not self.normalize_before:
            output = self.norm2(output)
        return output, self_attn


class TransformerEncoder(nn.Module):
    def __init__(self, num_layers, nhead, d_ffn, d_model, dropout, activation, normalize_before):
        super().__init__()
        self.layers = nn.ModuleList([
            TransformerEncoderLayer(d_ffn, nhead, d_model, dropout, activation, normalize_before)
            for _ in range(num_layers)])
        self.norm = LayerNorm(d_model, eps=1e-6)

    def forward(self, src, src_mask=None, src_key_padding_mask=None, pos_embs=None):
        output = src
        attention_lst = []
        for enc_layer in self.layers:
            output, attention = enc_layer(output, src_mask=src_mask,
                                          src_key_padding_mask=src_key_padding_mask,
                                          pos_embs=pos_embs)
            attention_lst.append(attention)
        return self.norm(output), attention_lst


# --------------------------------------------------------------------------
# Decoder side (the consumer right behind the path):
# speechbrain.lobes.models.transformer.Transformer.{TransformerDecoderLayer, TransformerDecoder,
# NormalizedEmbedding, get_lookahead_mask, get_key_padding_mask}, as reached from
# /root/reference/stac-st/modules/TransformerMultiTask.py:185-209 (forward) and :234-271 (decode).
# Same pinning status as the encoder classes: SpeechBrain's arithmetic restated from memory of ~v0.5.14
# (attribute names included: the cross-attention module really is spelled ``mutihead_attn`` there), the
# in-repo glue pinned by the reference's own decode()/forward() (tests/golden/make_decoder_golden.py).
# --------------------------------------------------------------------------
class _Embedding(nn.Module):
    """speechbrain/nnet/embedding.py::Embedding (wraps nn.Embedding as ``.Embedding``, padding_idx = blank_id)."""

    def __init__(self, num_embeddings, embedding_dim, blank_id=0):
        super().__init__()
        self.Embedding = nn.Embedding(num_embeddings, embedding_dim, padding_idx=blank_id)

    def forward(self, x):
        return self.Embedding(x.long())


class NormalizedEmbedding(nn.Module):
    """Embedding scaled by sqrt(d_model) (Transformer.py::NormalizedEmbedding)."""

    def __init__(self, d_model, vocab):
        super().__init__()
        self.emb = _Embedding(num_embeddings=vocab, embedding_dim=d_model, blank_id=0)
        self.d_model = d_model

    def forward(self, x):
        return self.emb(x) * math.sqrt(self.d_model)


def get_lookahead_mask(padded_input):
    """Float mask [L, L]: 0 on and below the diagonal, -inf above (Transformer.py::get_lookahead_mask)."""
    seq_len = padded_input.shape[1]
    mask = (torch.triu(torch.ones((seq_len, seq_len), device=padded_input.device)) == 1).transpose(0, 1)
    mask = mask.float().masked_fill(mask == 0, float("-inf")).masked_fill(mask == 1, float(0.0))
    return mask.detach().to(padded_input.device)


def get_key_padding_mask(padded_input, pad_idx):
    """Bool mask [B, L], True where the token is padding (Transformer.py::get_key_padding_mask)."""
    if len(padded_input.shape) == 4:
        bz, time, ch1, ch2 = padded_input.shape
        padded_input = padded_input.reshape(bz, time, ch1 * ch2)
    key_padded_mask = padded_input.eq(pad_idx).to(padded_input.device)
    return key_padded_mask.detach()


class TransformerDecoderLayer(nn.Module):
    def __init__(self, d_ffn, nhead, d_model, dropout, activation, normalize_before):
        super().__init__()
        self.self_attn = MultiheadAttention(nhead=nhead, d_model=d_model, dropout=dropout)
        self.mutihead_attn = MultiheadAttention(nhead=nhead, d_model=d_model, dropout=dropout)
        self.pos_ffn = PositionalwiseFeedForward(d_ffn=d_ffn, input_size=d_model,
                                                 dropout=dropout, activation=activation)
        self.norm1 = LayerNorm(d_model, eps=1e-6)
        self.norm2 = LayerNorm(d_model, eps=1e-6)
        self.norm3 = LayerNorm(d_model, eps=1e-6)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.normalize_before = normalize_before

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, pos_embs_tgt=None, pos_embs_src=None):
        tgt1 = self.norm1(tgt) if self.normalize_before else tgt
        tgt2, self_attn = self.self_attn(tgt1, tgt1, tgt1, attn_mask=tgt_mask,
                                         key_padding_mask=tgt_key_padding_mask)
        tgt = tgt + self.dropout1(tgt2)
        if not self.normalize_before:
            tgt = self.norm1(tgt)
        tgt1 = self.norm2(tgt) if self.normalize_before else tgt
        tgt2, multihead_attention = self.mutihead_attn(tgt1, memory, memory, attn_mask=memory_mask,
                                                       key_padding_mask=memory_key_padding_mask)
        tgt = tgt + self.dropout2(tgt2)
        if not self.normalize_before:
            tgt = self.norm2(tgt)
        tgt1 = self.norm3(tgt) if self.normalize_before else tgt
        tgt2 = self.pos_ffn(tgt1)
        tgt = tgt + self.dropout3(tgt2)
        if not self.normalize_before:
            tgt = self.norm3(tgt)
        return tgt, self_attn, multihead_attention


class TransformerDecoder(nn.Module):
    def __init__(self, num_layers, nhead, d_ffn, d_model, dropout, activation, normalize_before):
        super().__init__()
        self.layers = nn.ModuleList([
            TransformerDecoderLayer(d_ffn, nhead, d_model, dropout, activation, normalize_before)
            for _ in range(num_layers)])
        self.norm = LayerNorm(d_model, eps=1e-6)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, pos_embs_tgt=None, pos_embs_src=None):
        output = tgt
        self_attns, multihead_attns = [], []
        for dec_layer in self.layers:
            output, self_attn, multihead_attn = dec_layer(
                output, memory, tgt_mask=tgt_mask, memory_mask=memory_mask,
                tgt_key_padding_mask=tgt_key_padding_mask, memory_key_padding_mask=memory_key_padding_mask,
                pos_embs_tgt=pos_embs_tgt, pos_embs_src=pos_embs_src)
            self_attns.append(self_attn)
            multihead_attns.append(multihead_attn)
        return self.norm(output), self_attns, multihead_attns


def length_to_mask(length, max_len=None, dtype=None):
    """speechbrain/dataio/dataio.py::length_to_mask (the mask takes the dtype of `length` unless one is given, which
    is what lets the reference write ``1 - length_to_mask(enc_len)``, TransformerMultiTask.py:250)."""
    if max_len is None:
        max_len = length.max().long().item()
    mask = torch.arange(max_len, device=length.device, dtype=length.dtype).expand(
        len(length), max_len) < length.unsqueeze(1)
    return torch.as_tensor(mask, dtype=length.dtype if dtype is None else dtype, device=length.device)


class TransformerMultiTask(nn.Module):
    """Encoder half of /root/reference/stac-st/modules/TransformerMultiTask.py.

    ``encode``   follows :273-309 (mask ``j > floor(wav_len*T)``).
    ``forward_encoder`` follows the encoder half of ``forward`` :144-183 with
    ``make_masks`` :211-232 (mask ``~length_to_mask(round(wav_len*T))``).
    ``forward`` (decoder half :185-209) and ``decode`` :234-271 run SpeechBrain's TransformerDecoder over the
    encoder output.
    ``_init_params`` :311-314 re-initialises every dim>1 parameter with xavier_normal_.  The decoder and the target
    embedding are created and initialised under a forked random stream (after, not between, the encoder-side
    modules) so that the encoder-side weights drawn for a given seed - which the committed golden files depend on -
    are the same with and without a decoder; the distribution is the reference's.
    """

    def __init__(self, tgt_vocab, input_size, d_model=512, nhead=8,
                 num_encoder_layers=6, num_decoder_layers=6, d_ffn=2048,
                 dropout=0.1, activation=nn.ReLU, normalize_before=False,
                 max_length=2500, **_unused):
        super().__init__()
        self.positional_encoding = PositionalEncoding(d_model, max_length)
        self.encoder = TransformerEncoder(num_encoder_layers, nhead, d_ffn, d_model,
                                          dropout, activation, normalize_before)
        self.custom_src_module = _Layers(
            Linear(input_size=input_size, n_neurons=d_model, bias=True, combine_dims=False),
            nn.Dropout(dropout))
        self._init_params()
        self.decoder = None
        self.custom_tgt_module = None
        if num_decoder_layers > 0:
            with torch.random.fork_rng(devices=[]):
                torch.manual_seed(torch.initial_seed() + 1)
                self.decoder = TransformerDecoder(num_decoder_layers, nhead, d_ffn, d_model, dropout, activation,
                                                  normalize_before)
                self.custom_tgt_module = _Layers(NormalizedEmbedding(d_model, tgt_vocab))
                for m in (self.decoder, self.custom_tgt_module):
                    for p in m.parameters():
                        if p.dim() > 1:
                            nn.init.xavier_normal_(p)

    def _init_params(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_normal_(p)

    def forward(self, src, tgt, wav_len=None, pad_idx=0):
        """:144-209: encoder with the make_masks rule, then the decoder over the whole target."""
        encoder_out = self.forward_encoder(src, wav_len)
        src_key_padding_mask = None
        if wav_len is not None:
            abs_len = torch.round(wav_len * encoder_out.shape[1])
            src_key_padding_mask = ~length_to_mask(abs_len, max_len=encoder_out.shape[1]).bool()
        tgt_key_padding_mask = get_key_padding_mask(tgt, pad_idx=pad_idx)
        tgt_mask = get_lookahead_mask(tgt)
        tgt = self.custom_tgt_module(tgt)
        tgt = tgt + self.positional_encoding(tgt)
        decoder_out, _, _ = self.decoder(tgt=tgt, memory=encoder_out, memory_mask=None, tgt_mask=tgt_mask,
                                         tgt_key_padding_mask=tgt_key_padding_mask,
                                         memory_key_padding_mask=src_key_padding_mask)
        return encoder_out, decoder_out

    @torch.no_grad()
    def decode(self, tgt, encoder_out, enc_len=None):
        """:234-271: one decoding step = the whole decoder over the whole prefix."""
        tgt_mask = get_lookahead_mask(tgt)
        src_key_padding_mask = None
        if enc_len is not None:
            src_key_padding_mask = (1 - length_to_mask(enc_len)).bool()
        tgt = self.custom_tgt_module(tgt)
        tgt = tgt + self.positional_encoding(tgt)
        prediction, self_attns, multihead_attns = self.decoder(
            tgt, encoder_out, tgt_mask=tgt_mask, memory_key_padding_mask=src_key_padding_mask)
        return prediction, multihead_attns[-1]

    def _embed(self, src):
        if src.dim() == 4:
            bz, t, ch1, ch2 = src.shape
            src = src.reshape(bz, t, ch1 * ch2)
        return src

    def encode(self, src, wav_len=None):
        src = self._embed(src)
        src_key_padding_mask = None
        if wav_len is not None:
            abs_len = torch.floor(wav_len * src.shape[1])
            src_key_padding_mask = (
                torch.arange(src.shape[1])[None, :].to(abs_len) > abs_len[:, None])
        src = self.custom_src_module(src)
        src = src + self.positional_encoding(src)
        encoder_out, _ = self.encoder(src=src, src_key_padding_mask=src_key_padding_mask)
        return encoder_out

    def forward_encoder(self, src, wav_len=None):
        src = self._embed(src)
        src_key_padding_mask = None
        if wav_len is not None:
            abs_len = torch.round(wav_len * src.shape[1])
            src_key_padding_mask = ~length_to_mask(abs_len, max_len=src.shape[1]).bool()
        src = self.custom_src_module(src)
        src = src + self.positional_encoding(src)
        encoder_out, _ = self.encoder(src=src, src_mask=None,
                                      src_key_padding_mask=src_key_padding_mask)
        return encoder_out


class _Layers(nn.Module):
    """Stand-in for speechbrain.nnet.containers.ModuleList (children under ``.layers``)."""

    def __init__(self, *layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class EncoderWrapper(nn.Module):
    """/root/reference/stac-st/modules/TransformerMultiTask.py:317-349."""

    def __init__(self, transformer):
        super().__init__()
        self.transformer = transformer

    def forward(self, x, wav_lens=None):
        return self.transformer.encode(x, wav_lens)


# --------------------------------------------------------------------------
# The object graph the reference yaml instantiates, and its call sequence.
# --------------------------------------------------------------------------
def build_reference_modules(size="S", vocab=5000, n_mels=80, seed=8886, **overrides):
    """Instantiate what transformer_multitask.yaml:173-210,253-254,299-302 builds
    (encoder side), seeded like the yaml (seed 8886, :22-23)."""
    cfg = dict(MODEL_SIZES[size]) if isinstance(size, str) else dict(size)
    cfg.update(overrides)
    torch.manual_seed(seed)
    mods = {
        "compute_features": Fbank(sample_rate=16000, n_fft=400, n_mels=n_mels),
        "normalize": InputNormalization(norm_type="global", update_until_epoch=4),
        "CNN": ConvolutionFrontEnd(input_shape=(8, 10, n_mels), num_blocks=2,
                                   num_layers_per_block=1, out_channels=(256, 256),
                                   kernel_sizes=(3, 3), strides=(2, 2),
                                   residuals=(False, False)),
        "Transformer": TransformerMultiTask(
            input_size=(((n_mels - 1) // 2 + 1) - 1) // 2 * 256 + 256, tgt_vocab=vocab,
            d_model=cfg["d_model"], nhead=cfg["nhead"],
            num_encoder_layers=cfg["num_encoder_layers"], num_decoder_layers=6,
            d_ffn=cfg["d_ffn"], dropout=0.1, activation=nn.GELU,
            normalize_before=True, causal=False),
        "ctc_lin": Linear(input_size=cfg["d_model"], n_neurons=vocab),
        "log_softmax": nn.LogSoftmax(dim=-1),
    }
    for m in mods.values():
        m.eval()
    return mods


@torch.no_grad()
def reference_compute_forward(mods, wavs, wav_lens, train_mask=False, stages=None):
    """/root/reference/stac-st/inference.py:95-107 (train_mask=False) or
    /root/reference/stac-st/train_multitask.py:59-78 encoder half (train_mask=True).
    Returns a dict of every stage boundary tensor."""
    out = {}
    feats = mods["compute_features"](wavs)
    out["fbank"] = feats
    feats = mods["normalize"](feats, wav_lens)
    out["feats"] = feats
    src = mods["CNN"](feats)
    out["cnn"] = src
    if stages == "frontend":
        return out
    tr = mods["Transformer"]
    enc_out = tr.forward_encoder(src, wav_lens) if train_mask else tr.encode(src, wav_lens)
    out["enc_out"] = enc_out
    logits = mods["ctc_lin"](enc_out)
    out["logits"] = logits
    out["p_ctc"] = mods["log_softmax"](logits)
    return out
