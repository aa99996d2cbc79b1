"""CPU oracle for the population hot path of keurfonluu/stochopy (v2.3.0).

TEST INFRASTRUCTURE ONLY.  This package is a numpy restatement of the
reference's per-generation algorithms (objective evaluation, DE / PSO / CPSO
updates and selection, CMA-ES / VD-CMA sampling and updates, NA resampling).
It exists to check the CUDA path in ``stochopy_b200`` and to time the
reference's CPU algorithm beside it.  Only ``tests/``, ``__graft_entry__.smoke``
and ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``) may import it;
nothing under ``stochopy_b200/`` does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
here against (a) the 24 known-answer cases of the reference's own
``tests/test_optimize.py`` + the README run, (b) ``tests/test_factory.py``'s
seven objective values and (c) step-level fixtures in ``tests/golden/`` that
were produced by importing the real reference in the dev container
(``tests/golden/make_golden.py``).

Every random number is an explicit input of a step function.  Two stream
providers produce them:

* ``streams.MTStream`` draws from numpy's legacy MT19937 generator in exactly
  the order the reference consumes it, which is
  what makes the reference's golden vectors reproducible;
* ``streams.PhiloxStream`` reproduces, bit for bit, the counter-based
  Philox4x32-10 draws of the CUDA kernels, so the device's own random mode
  can be checked against the same step functions.
"""
