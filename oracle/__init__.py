"""CPU oracle for the VAE ELBO train-step hot path of virajmehta/vae-training.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import or execute anything under
``oracle/``.  The product path (``vae_training_b200``) never imports it and fails
loudly when its CUDA library is missing.

What is restated here (citations are into the read-only reference tree):

* ``jax_random.py``  -- third-party ``jax.random`` (threefry2x32, ``split``,
  ``uniform``, ``normal``) as used at ``model.py:29,140,154,208,227``,
  ``vae.py:56,125`` and ``datasets.py:35,51,76,173,185,237,242``.  jax/jaxlib are
  un-vendored and un-pinned (``README.md:8-13``); the algorithm is restated from
  the published Random123 / jax.random definition.
* ``datasets.py``    -- ``datasets.py:34-52,55-98,163-228,230-279``.
* ``vae_step.py``    -- ``networks.py:26-44,59-113`` + legacy ``flax.optim.Adam``
  (``vae.py:113``), with the reverse-mode gradient derived by hand.
* ``trainer.py``     -- the RNG / event schedule of ``model.py:207-228`` and
  ``vae.py:123-141,191-201``.
* ``c/vae_oracle.c`` -- the same step and loop in plain C (the timed CPU baseline).

PARITY PINNING STATUS
---------------------
The reference has no tests, golden vectors or fixtures of any kind and cannot be
executed here or on the GPU box (no jax / flax wheels, no network).  What pins the
oracle:

* integer path (threefry2x32 bits, ``split``, key chains): pinned by the three
  Random123 known-answer vectors and the documented ``split(PRNGKey(0))`` value --
  see ``tests/test_oracle_rng.py``;
* float path of ``normal``: pinned only on ``uniform(PRNGKey(0))``,
  ``normal(PRNGKey(0))`` and ``normal(PRNGKey(42))``; XLA-CPU's ``log1p`` codegen
  is not reproducible here, so beyond those values the claim is "<= 1 ulp" --
  **parity unpinned**;
* step arithmetic (forward / backward / Adam): checked against torch autograd in
  float64 and a torch Adam-style reference; **parity unpinned** against the real
  flax/XLA executable.
"""
