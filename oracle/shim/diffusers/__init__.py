"""TEST INFRASTRUCTURE ONLY. Minimal restatement of the parts of diffusers 0.21.4 (pinned by the reference's
THIRD-PARTY:31; not installed here) that the reference's hot path imports, so that /root/reference's own modules can be
imported unchanged in this container to pin oracle/insv2v_oracle.py and mint tests/golden. Never imported by the
product."""
from .schedulers import DDIMScheduler, DDPMScheduler, PNDMScheduler  # noqa: F401
