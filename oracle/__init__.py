"""CPU oracle for the nmrfit objective-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nmrfit_b200/`` may import this
package; the only permitted importers are ``tests/`` (including its fuzz drivers
``tests/fuzz_*.py``), ``__graft_entry__.smoke()`` and the CPU legs of the benchmarks
(``cpu_baseline`` / ``--impl reference`` in ``bench.py``; the quadrature timing in
``tools/bench_curves.py``, config 5's counterpart), and there only as the checker or
as the CPU arm being timed, never as the product path.

Parity status
-------------
* ``nmrfit_oracle`` (ps2 / voigt / objective / laplace1d / weights /
  generate_result / Kramers-Kronig): PINNED.  The reference ships no tests or
  golden vectors, so the pins are outputs of the unmodified
  reference itself, imported from /root/reference in the build container by
  ``tests/golden/make_golden.py`` and committed under ``tests/golden/``.
* ``nmrfit_oracle.auto_peaks`` (AutoPeakSelector.find_peaks, utils.py:670-783): PINNED against peak lists
  the reference's own selector produced (``tests/golden/peaks_*.npz``), with ``peakutils_oracle.baseline``
  standing in for the one third-party call inside it.
* ``peakutils_oracle`` (peakutils.baseline): PARITY UNPINNED.  peakutils is listed in the reference's
  requirements.txt:4 without a version and is absent from this image; the iterative polynomial baseline is
  restated from its published algorithm and anchored on the reference's call sites (utils.py:719, 766).
* ``ref_loader``: imports the UNMODIFIED reference package (``baseline/_ref``, else ``/root/reference``) with the
  shims SURVEY.md Appendix A lists (numpy aliases removed since 1.24, matplotlib / nmrglue / peakutils absent,
  ``pyswarm`` bound to ``pso_oracle``); used by the golden generator, the drop-in test and the CPU arm of bench.py.
* ``pso_oracle`` (pyswarm.pso): PARITY UNPINNED.  pyswarm is a third-party,
  un-vendored, un-pinned dependency of the reference (README.md:13-17 installs
  git master of tisimst/pyswarm; it is absent from requirements.txt and from this
  image).  The loop is restated from its published algorithm and anchored only on
  the reference's call site (nmrfit/utils.py:176-182).
"""
