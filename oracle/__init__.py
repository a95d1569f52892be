"""CPU oracle for the masked selective-scan hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``actalker_b200/`` may import this package.  The only allowed
importers are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, and there only as the checker or
as the timed CPU baseline, never as the thing shipped.

Pin status:
  * operator level (`selective_scan_ref`): PARITY UNPINNED by upstream tests —
    the arithmetic lives in the third-party package mamba-ssm==1.2.0.post1
    (reference pins: environment.yaml:43, install_actalker.sh:93) which is not
    vendored under /root/reference and not installable here.  The restatement
    follows the package's published semantics and is
    cross-checked against the independent statement of the same recurrence in
    transformers' `MambaMixer.slow_forward` (tests/test_oracle.py) and, on the GPU
    box, the CUDA path is compared with the mamba-ssm-derived selective-scan
    kernel that ships in vLLM (tests/test_gpu_parity.py, test_gpu_fullsize.py).
  * mask level (`downsample`): restated from diffusers==0.29.2
    (requirements.txt:10), also absent; pinned by known-answer cases.
  * layer level (`SS2D_cond_v10_ref`, `SS2D_Unit_ref`): PINNED — the real
    reference classes from /root/reference/src/models/base/mamba_layer.py are
    executed in the build container (third-party imports stubbed with the two
    restatements above) by tests/golden/make_golden.py and the outputs are
    committed under tests/golden/.
"""
from .selective_scan_ref import selective_scan_ref  # noqa: F401
from .mask_downsample import downsample, mask_to_index  # noqa: F401
from .mamba_layer_ref import HSCANS_dynamic_ref, SS2D_Unit_ref, SS2D_cond_v10_ref  # noqa: F401
