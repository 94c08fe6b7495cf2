"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU fp32 restatement of the medimgen hot path (strided DiffusionModelUNet / AutoencoderKL blocks
and the DDPMScheduler add_noise/step loop). Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it, and only as the checker or the
timed CPU baseline. Nothing under `medical_image_generation_b200/` imports from here.

Pinning status (see DESIGN.md "Oracle"):
  * model blocks (rows a1-a10): pinned against the UNMODIFIED reference modules executed in the
    build container under the 4-symbol MONAI shim (oracle/shim); outputs committed as
    tests/golden/*.pt by oracle/gen_golden.py.
  * scheduler / inferers (rows a11-a12): arithmetic lives in the absent, unpinned third-party
    package `monai-generative` -> PARITY UNPINNED by the reference; restated from its published
    algorithm and cross-checked by closed-form identities only.
"""
