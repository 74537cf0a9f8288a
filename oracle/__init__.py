"""CPU oracle for the NMF spectrogram-inpainting hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it, and only as the checker (or as the thing timed for the CPU baseline), never as a
fallback for the CUDA path.

Three layers, strongest pin first:

* ``oracle.ref_loader``  -- imports the *unmodified* reference scripts from ``/root/reference``
  (only possible in the build container; used to generate ``tests/golden/*``).
* ``oracle.libcalls``    -- the reference's ~15 lines of numpy glue per script restated around the
  *same* third-party calls it makes (``scipy.signal.stft/istft``, ``sklearn.decomposition.NMF``).
  This is the CPU baseline that ``bench.py`` times (kind "port": glue restated, arithmetic
  executed by the very same scipy/sklearn/OpenBLAS wheels the reference runs on).
* ``oracle.restate``     -- first-principles restatement of what those wheels compute (framing,
  window, rFFT scaling, column mask predicate, imputation, MT19937 initial factors, the
  coordinate-descent loop via ``oracle/c/oracle_c.c``, recombination, overlap-add), exposing every
  intermediate so a failing CUDA stage can be located.

Pinning status: the reference ships no tests.  The oracle is pinned against (a) the reference's
shipped output WAVs ``demo_assets/part0/nmf_{original,corrupted,restored}.wav`` and
``demo_assets/part2/fixed_nmf_gap.wav`` (<= 1 int16 LSB) and (b) outputs of the reference scripts
themselves imported in the build container (``tests/golden/make_golden.py``).
"""
