"""insv2v_b200 — B200-native (sm_100a) implementation of the InsV2V denoising hot path:
UNet3DConditionModel.forward, AutoencoderKL.decode and the optical-flow warp, behind the reference's own
class/function names. Host side is thin Python over the C ABI in include/ivv.h."""
__version__ = "0.1.0"
