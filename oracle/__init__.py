"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (plain PyTorch fp32 + numpy integer code) of the MMGT stage-2 denoising hot
path, used as the *checker* for the CUDA path in ``mmgt_b200/``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  The product path never imports it and fails loudly without the CUDA library.

Parity pin: the reference ships NO tests / golden vectors, so the oracle is pinned
against outputs of the reference's own modules, imported unchanged from ``/root/reference``
on top of the ``ref_shim/diffusers`` stand-in (``make_golden.py`` -> ``tests/golden/*.npz``).
The third-party layer (diffusers==0.24.0, requirements.txt:36) is restated from its published
behaviour and could not be diffed against the real package here => for that layer only the
status is "parity unpinned" (see DESIGN.md section Oracle).
"""
