"""CPU oracle for the dueling double-DQN train step + replay path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker or as the timed CPU
baseline -- never as part of the CUDA path.

Parity status
-------------
* Replay half (``replay_oracle``): PINNED.  Checked bit-for-bit against outputs of the
  reference's own ``General/Base/replay_buffer.py`` (``ReplayBuffer.add``, numba
  ``sample_batch``) run in the build container; vectors frozen in ``tests/golden/`` by
  ``oracle/make_golden.py``.
* Train-step half (``dqn_oracle``): PARITY UNPINNED.  jax / dm-haiku / optax are not
  installable here (no network, no wheels), the reference ships no tests and no
  input/output vectors, so the arithmetic of ``q_learning_functions.py`` / ``dddqn.py`` /
  optax ``adam``/``adamw``/``huber_loss`` is a restatement from source + published
  library semantics.  It is cross-validated by an independent torch-autograd derivation
  (``tests/test_oracle_autograd.py``) and anchored on the reference's only golden
  artefacts, ``Test/lunar_lander/{params,opt_state}.pickle`` (tree names, layouts, theta_0).
"""
