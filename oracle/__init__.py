"""CPU oracles -- TEST INFRASTRUCTURE ONLY.

Nothing under ``harmonypy_b200/`` imports this package (tests/test_repo_contracts.py enforces it); only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

  harmony_oracle.py      the Harmony inner loop (harmonypy/harmony.py:366-569), pinned to the real reference by
                         tests/golden/ fixtures and by live differential tests (fp64: 1e-9)
  lisi_oracle.py         compute_lisi (harmonypy/lisi.py), pinned to the reference's known-answer test
  kmeans_init_oracle.py  NumPy restatement of the optional device-side centroid initialisation (no reference
                         counterpart: sklearn's random stream cannot be reproduced; inertia on par with sklearn)
  device_perm.py         NumPy mirror of the engine's device-side block permutation, so that perm_mode="device"
                         runs can be replayed through harmony_oracle.py
"""
