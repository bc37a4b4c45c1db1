"""CPU oracle for the STAC-ST encoder-side inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and there only as the checker
(or the timed CPU baseline), never as the thing shipped.

PARITY UNPINNED: the arithmetic of this path lives in SpeechBrain (~v0.5.14,
``/root/reference/README.md:46-50``), which is neither vendored under
``/root/reference`` nor installable here, and the reference ships no tests or
golden vectors.  The restatement in
``oracle/speechbrain_path.py`` follows the published SpeechBrain source from
memory of that release; it is cross-checked against independent formulations
(numpy DFT, hand-written attention, hand-written LayerNorm/conv) in
``tests/test_oracle.py`` and against the three constants the reference does
pin (5120-wide CNN output, 25 Hz frame rate, 2500-entry PE table).

What IS pinned by the reference itself: the in-repo glue.  ``tests/golden/make_glue_golden.py``
imports ``/root/reference/stac-st/modules/TransformerMultiTask.py`` unmodified (its seven
SpeechBrain imports served by a stub package made of this oracle's classes) and records what the
reference's own ``encode()`` / ``forward()`` / ``make_masks()`` / ``EncoderWrapper`` return on seeded
inputs in ``tests/golden/glue_reference.npz``; ``tests/test_oracle.py`` holds the restated
``TransformerMultiTask`` to those vectors at 1e-6 and ``tests/test_gpu_glue_reference.py`` holds the
CUDA encoder to them at the north-star tolerances.

The two consumers right behind the path are pinned the same way: ``tests/golden/make_decoder_golden.py``
runs the reference's own ``decode()`` / ``forward()`` (decoder half) over the restated TransformerDecoder
(``tests/golden/decoder_reference.npz``), and ``tests/golden/make_turns_golden.py`` executes the reference's own
``append_speaker_turns`` body, cut out of inference.py with ``ast`` (``tests/golden/turns_reference.json``), which pins
``oracle/turns.py`` completely (that function has no SpeechBrain arithmetic in it).
"""
from .speechbrain_path import (  # noqa: F401
    Fbank,
    InputNormalization,
    ConvolutionFrontEnd,
    TransformerMultiTask,
    EncoderWrapper,
    Linear,
    build_reference_modules,
    reference_compute_forward,
    MODEL_SIZES,
)
