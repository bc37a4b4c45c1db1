"""Drop-ins for ``hparams.compute_features`` and ``modules.normalize``.

They take the constructor arguments the reference yaml passes
(/root/reference/stac-st/hparams/transformer_multitask.yaml:299-302 and :208-210) and keep
SpeechBrain's call signatures (/root/reference/stac-st/inference.py:95-96,
/root/reference/stac-st/train_multitask.py:59-61); the arithmetic runs in libstac_b200.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._lib import StacB200Error


class Fbank(nn.Module):
    """``speechbrain.lobes.features.Fbank`` replacement (log-mel filterbank features).

    Only the configuration the reference uses is implemented on the GPU: 16 kHz, n_fft 400,
    25 ms hamming window, 10 ms hop, 80 triangular mel filters 0-8 kHz, no deltas / context.
    ``top_db_per_utterance`` mirrors SpeechBrain's per-sequence top-dB clamp (False = the older
    batch-global clamp).  Like SpeechBrain's, the module has no parameters or buffers.
    """

    def __init__(self, deltas=False, context=False, requires_grad=False, sample_rate=16000, f_min=0,
                 f_max=None, n_fft=400, n_mels=40, filter_shape="triangular", param_change_factor=1.0,
                 param_rand_factor=0.0, left_frames=5, right_frames=5, win_length=25, hop_length=10,
                 top_db_per_utterance=True):
        super().__init__()
        f_max = sample_rate / 2 if f_max is None else f_max
        supported = (not deltas and not context and not requires_grad and sample_rate == 16000 and f_min == 0
                     and f_max == 8000 and n_fft == 400 and n_mels == 80 and filter_shape == "triangular"
                     and win_length == 25 and hop_length == 10)
        if not supported:
            raise StacB200Error("stac_b200 Fbank implements the STAC-ST configuration only "
                                "(sample_rate=16000, n_fft=400, n_mels=80, triangular, no deltas/context)")
        self.top_db = 80.0
        self.top_db_per_utterance = top_db_per_utterance
        self._tables = None

    def tables(self, device):
        if self._tables is None or self._tables.device != device:
            self._tables = ops.build_fbank_tables(device)
        return self._tables

    def tc_tables(self, device):
        """Constants of the tensor-core STFT kernel (used by the fused bf16 pipeline)."""
        if getattr(self, "_tc_tables", None) is None or self._tc_tables[0].device != device:
            self._tc_tables = ops.build_fbank_tc_tables(device)
        return self._tc_tables

    @torch.no_grad()
    def forward(self, wav):
        return ops.fbank(wav.float(), self.tables(wav.device), self.top_db, self.top_db_per_utterance)


class InputNormalization(nn.Module):
    """``speechbrain.processing.features.InputNormalization`` replacement (norm_type="global").

    Eval-mode forward is ``(x - glob_mean) / glob_std`` on the GPU.  The running statistics are
    plain attributes saved/restored through ``_save`` / ``_load`` exactly like SpeechBrain's
    ``normalizer.ckpt`` (dict keys count/glob_mean/glob_std/spk_dict_*).  Updating the statistics
    (train mode, ``train_multitask.py:60-61``): the per-utterance mean / std over the valid frames come from
    ``stac_utt_mean_std`` on the device, the running-average update of the 80 global values is SpeechBrain's code on the
    host copies.  ``calibrate`` computes them once from a batch with
    SpeechBrain's formula.
    """

    def __init__(self, mean_norm=True, std_norm=True, norm_type="global", avg_factor=None,
                 requires_grad=False, update_until_epoch=3):
        super().__init__()
        if norm_type != "global" or not mean_norm or not std_norm:
            raise StacB200Error("stac_b200 InputNormalization implements norm_type='global' with mean and std")
        self.mean_norm, self.std_norm, self.norm_type = mean_norm, std_norm, norm_type
        self.avg_factor = avg_factor
        self.update_until_epoch = update_until_epoch
        self.glob_mean = torch.tensor([0])
        self.glob_std = torch.tensor([0])
        self.spk_dict_mean, self.spk_dict_std, self.spk_dict_count = {}, {}, {}
        self.weight = 1.0
        self.count = 0
        self.eps = 1e-10
        self._dev_stats = None

    @torch.no_grad()
    def forward(self, x, lengths=None, spk_ids=torch.tensor([]), epoch=0):
        if self.training:
            if lengths is None:
                raise StacB200Error("InputNormalization needs the relative lengths in training mode")
            self._update_statistics(x, lengths, epoch)
        mean, std = self.device_stats(x.device, x.shape[-1])
        return ops.input_norm(x.float(), mean, std)

    def _update_statistics(self, x, lengths, epoch):
        """SpeechBrain's train-mode step: per-utterance statistics (device), batch average and running update (host)."""
        x = x.float().contiguous()
        b, t, f = x.shape
        wl = lengths.to(device=x.device, dtype=torch.float32).contiguous()
        means = torch.empty(b, f, device=x.device, dtype=torch.float32)
        stds = torch.empty(b, f, device=x.device, dtype=torch.float32)
        ops._call("stac_utt_mean_std", ops.ptr(x, torch.float32), ops.ptr(wl, torch.float32), b, t, f, float(self.eps),
                  ops.ptr(means), ops.ptr(stds), ops.stream())
        current_mean = torch.mean(means.cpu(), dim=0)
        current_std = torch.mean(stds.cpu(), dim=0)
        if self.count == 0:
            self.glob_mean, self.glob_std = current_mean, current_std
        elif epoch < self.update_until_epoch:
            self.weight = 1 / (self.count + 1) if self.avg_factor is None else self.avg_factor
            self.glob_mean = (1 - self.weight) * self.glob_mean.cpu() + self.weight * current_mean
            self.glob_std = (1 - self.weight) * self.glob_std.cpu() + self.weight * current_std
        self.count = self.count + 1
        self._dev_stats = None

    def device_stats(self, device, n_mels):
        if self.glob_mean.numel() != n_mels:
            raise StacB200Error("InputNormalization has no statistics yet: load normalizer.ckpt or call calibrate()")
        key = (device, self.glob_mean.data_ptr(), self.glob_std.data_ptr(), self.count)
        if self._dev_stats is None or self._dev_stats[0] != key:
            self._dev_stats = (key, self.glob_mean.detach().to(device, torch.float32).contiguous(),
                               self.glob_std.detach().to(device, torch.float32).contiguous())
        return self._dev_stats[1], self._dev_stats[2]

    @torch.no_grad()
    def calibrate(self, feats, lengths):
        """One SpeechBrain train-mode statistics step (count == 0 branch) on an un-normalised batch."""
        means, stds = [], []
        for i in range(feats.shape[0]):
            n = int(torch.round(lengths[i] * feats.shape[1]))
            seg = feats[i, :n].float()
            means.append(seg.mean(0))
            stds.append(torch.max(seg.std(0), torch.full_like(seg[0], self.eps)))
        self.glob_mean = torch.stack(means).mean(0).cpu()
        self.glob_std = torch.stack(stds).mean(0).cpu()
        self.count = 1
        self._dev_stats = None

    # --- SpeechBrain checkpoint hooks (normalizer.ckpt) ---
    def _statistics_dict(self):
        return {"count": self.count, "glob_mean": self.glob_mean, "glob_std": self.glob_std,
                "spk_dict_mean": self.spk_dict_mean, "spk_dict_std": self.spk_dict_std,
                "spk_dict_count": self.spk_dict_count}

    def _load_statistics_dict(self, state):
        self.count = state["count"]
        self.glob_mean = state["glob_mean"]
        self.glob_std = state["glob_std"]
        self.spk_dict_mean = state.get("spk_dict_mean", {})
        self.spk_dict_std = state.get("spk_dict_std", {})
        self.spk_dict_count = state.get("spk_dict_count", {})
        self._dev_stats = None
        return state

    def _save(self, path):
        torch.save(self._statistics_dict(), path)

    def _load(self, path, end_of_epoch=False, device=None):
        del end_of_epoch
        self._load_statistics_dict(torch.load(path, map_location=device))
