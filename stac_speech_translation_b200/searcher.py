"""KV-cached ``forward_step`` for the reference's beam searcher.

The reference's ``S2SMultiTaskTransformerBeamSearch`` (/root/reference/stac-st/modules/mutitask_decoder.py:14-137) is
SpeechBrain's ``S2SBeamSearcher`` plus three small methods: ``reset_mem`` :101-103 (the ``[bos, source_lang,
target_lang]`` prefix), ``permute_mem`` :109-112 (beam re-ordering of the token memory) and ``forward_step`` :119-128,
which re-runs the whole decoder over the whole prefix at every step.  ``CachedStepMixin`` overrides exactly those three so
that the search - SpeechBrain's host-side control flow, untouched - drives ``decoder.DecoderCache`` instead:

    from modules.mutitask_decoder import S2SMultiTaskTransformerBeamSearch          # the reference class
    from stac_speech_translation_b200.searcher import CachedStepMixin

    class FastSearch(CachedStepMixin, S2SMultiTaskTransformerBeamSearch):
        pass
    # yaml: test_search: !new:<module>.FastSearch   (same arguments as transformer_inference.yaml:144-156)

What stays the reference's: the token memory that ``forward_step`` returns (the searcher reads the hypotheses from it),
the log-softmax over ``self.fc(pred) / self.temperature``, ``set_decoder_prefix_tokens``.  What changes: the decoder sees
one new token per step; the encoder states are not inflated x beam for the cross-attention (row r reads utterance
r // beam_size); the returned attention is that of the new position only ([rows, 1, frames] instead of [rows, length,
frames]; SpeechBrain uses it only for ``using_max_attn_shift`` / coverage penalties, which the reference leaves off).
"""
from __future__ import annotations

import torch

from ._lib import StacB200Error


def _update_mem(inp_tokens, memory):
    """mutitask_decoder.py:140-154."""
    if memory is None:
        return inp_tokens.unsqueeze(1)
    return torch.cat([memory, inp_tokens.unsqueeze(1)], dim=-1)


class CachedStepMixin:
    """Mix in front of the reference's ``S2SMultiTaskTransformerBeamSearch`` (see the module docstring).  Expects the
    attributes that class has: ``model`` (this package's TransformerMultiTask), ``fc``, ``softmax``, ``temperature``,
    ``bos_index``, ``decoder_input_tokens``, ``beam_size``, ``max_decode_ratio``.  The cache is sized from the search
    itself: SpeechBrain's searcher runs at most ``int(enc_states.shape[1] * max_decode_ratio)`` steps (the reference yamls
    set max_decode_ratio 1.0, i.e. up to 750 steps for a 30 s segment), bounded by the positional-encoding table."""

    def _cache_len(self, prefix_len: int, enc_frames: int) -> int:
        ratio = float(getattr(self, "max_decode_ratio", 1.0))
        want = prefix_len + int(enc_frames * ratio) + 1
        pe = getattr(getattr(self.model, "positional_encoding", None), "pe", None)
        return min(want, int(pe.shape[1])) if pe is not None else want

    def reset_mem(self, batch_size, device):
        self._kv_cache = None
        return torch.tensor([self.decoder_input_tokens] * batch_size).to(device)

    def permute_mem(self, memory, index):
        if getattr(self, "_kv_cache", None) is not None:
            self._kv_cache.reorder(index)
        return torch.index_select(memory, dim=0, index=index)

    def forward_step(self, inp_tokens, memory, enc_states, enc_lens):
        if not torch.all(inp_tokens == self.bos_index):
            memory = _update_mem(inp_tokens, memory)
        cache = getattr(self, "_kv_cache", None)
        if cache is None:
            rows = memory.shape[0]
            beam = int(getattr(self, "beam_size", 1))
            if rows % beam != 0 or enc_states.shape[0] != rows:
                raise StacB200Error("forward_step expects encoder states inflated to one row per hypothesis")
            # the searcher inflated the encoder states x beam (repeat_interleave): keep one copy per utterance
            # (decoder_precision = "bf16" on the searcher: the step's GEMMs on the tensor cores; decoder_graph = True: a
            # step is one CUDA-graph replay)
            cache = self.model.decoder_cache(enc_states[::beam].contiguous(), rows=rows,
                                             max_len=self._cache_len(memory.shape[1], enc_states.shape[1]),
                                             precision=getattr(self, "decoder_precision", "fp32"),
                                             graph=bool(getattr(self, "decoder_graph", False)))
            self._kv_cache = cache
            for t in range(memory.shape[1]):              # the [bos, source_lang, target_lang] prefix
                pred, attn = cache.step(memory[:, t].contiguous())
        else:
            pred, attn = cache.step(memory[:, -1].contiguous())
        prob_dist = self.softmax(self.fc(pred.unsqueeze(1)) / self.temperature)
        return prob_dist[:, -1, :], memory, attn.unsqueeze(1)
