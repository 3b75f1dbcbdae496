"""CPU oracle for the FA-VAE VQ-search + spectrum-loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``favae_b200/`` imports this package.
The only permitted users are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only
as the checker or as the timed CPU baseline, never as the product path.

Parity status (see DESIGN.md "Oracle"):

* ``vq_oracle``      -- PINNED against the reference's own ``models/l2_quantize.py``
                        run on CPU in the authoring container; fixtures under
                        ``tests/golden/vq_*.npz`` (made by ``oracle/make_golden.py``).
* ``blur_oracle``    -- PINNED against the reference's ``VQGANFCM._gaussian_blur``
                        (``models/vqgan_fcm.py:20-41``) and ``torchvision`` GaussianBlur
                        (``losses/vqgan_losses.py:35``); fixtures ``tests/golden/blur_*.npz``.
* ``wrappers_oracle``-- PINNED against ``losses/vqgan_losses.py`` (imported unmodified).
* ``ffl_oracle``     -- **PARITY UNPINNED**.  The arithmetic lives in the un-vendored
                        PyPI package ``focal-frequency-loss==0.3.0``
                        (``environment.yaml:139``), absent from /root/reference and not
                        installable offline.  The restatement follows the published
                        algorithm and is anchored on analytic known-answer tests
                        (``tests/test_oracle_ffl.py``) and on the reference call sites
                        ``favae_scripts/train_favae.py:313,318,326``.
"""
