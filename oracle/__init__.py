"""CPU oracle for the ToeplitzMatrices.jl FFT hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import or execute it, and there only as the
checker or as the timed CPU baseline -- never as the thing shipped.  The
product path (``toeplitzmatrices.jl_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Parity status: **parity unpinned** in the strict sense -- the reference
(ToeplitzMatrices.jl v0.8.5) ships no golden vectors or known-answer files
(``test/runtests.jl`` draws its inputs from Julia's version-specific RNG) and
neither ``julia`` nor FFTW exists in this image, so the reference itself
cannot be run here.  What pins the oracle instead (see
``tests/test_oracle_pinning.py``): every deterministic assertion of the
reference's own test-suite for this path, restated -- dense-product equality
at ``rtol = sqrt(eps)`` on the reference's coefficient recipes
(``test/runtests.jl:23-63,83-110,154-215``), the exact small cases
(``:76-77,156-161,797``), the spectrum convention ``fft(C.vc) ==
factorize(C).vcvr_dft`` (``:543-564``) and the n=8 Levinson/Durbin/Trench
kernels against dense solves (``:112-151``).
"""
