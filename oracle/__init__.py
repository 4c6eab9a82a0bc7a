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
* Train-step half (``dqn_oracle``, ``agent_oracle``): PINNED TO THE REFERENCE'S OWN SOURCES, with the third-party
  primitives beneath them restated.  jax / dm-haiku / optax are not installable here (no network, no wheels) and the
  reference ships no tests or input/output vectors.  ``oracle/make_golden_train.py`` therefore imports
  ``General/QLearning/q_learning_functions.py`` and ``LunarLander/dddqn.py`` UNMODIFIED from ``/root/reference`` on top
  of ``oracle/ref_shims`` (minimal jax / haiku / optax / gym modules: ``jit`` = identity, ``grad`` = reverse-mode
  autograd in float32, ``hk.Linear`` = ``x @ w + b``, optax 0.1.x ``adam`` / ``adamw`` / ``huber_loss``; see its
  README) and freezes what the reference's closures compute -- ``preprocessing`` -> ``compute_q_targets`` ->
  ``train_step``, several steps with hard syncs, ``compute_action`` -- into ``tests/golden/train_ref_*.npz``.
  ``tests/test_oracle_train_golden.py`` holds the oracle to those vectors; ``tests/test_gpu_train_golden.py`` holds the
  kernels to them directly; ``oracle/make_golden_agent_step.py`` records the reference's own ``Agent._step()`` and
  ``ParamAgent.inject()`` + ``_step()`` entry points the same way (``tests/golden/agent_step_ref.npz``).  What remains restated (and is named as such) is the arithmetic of the absent libraries'
  primitives, not the reference's code.  It is additionally cross-validated by an independent torch-autograd derivation
  (``tests/test_oracle_autograd.py``) and anchored on ``Test/lunar_lander/{params,opt_state}.pickle`` (tree names,
  layouts, theta_0).
* Episode loop (``episode_oracle``): PINNED the same way: ``oracle/make_golden_episode.py`` runs the reference's own
  ``Agent.training()`` (``General/QLearning/q_agent.py`` unmodified, scripted environment whose rewards / dones do not
  depend on the actions) and records its control flow into ``tests/golden/episode_ref_*.npz``
  (``tests/test_oracle_episode_golden.py``, ``tests/test_gpu_episode.py``).  The epsilon-greedy DRAWS are a stated
  Philox convention (the reference's host RNGs are global and its sampler RNG cannot be seeded).
* Prioritized replay (``per_oracle``): self-specified -- the reference has none.
"""
