"""The InsV2V inference configuration (`configs/instruct_v2v_inference.yaml:22-89` of the reference) as Python dicts:
the `params:` blocks a caller passes to `UNet3DConditionModel(**UNET_PARAMS)` / `AutoencoderKL(**VAE_PARAMS)` — what
`misc_utils.model_utils.instantiate_from_config` does with the YAML."""

UNET_PARAMS = dict(
    in_channels=8, out_channels=4, act_fn="silu", attention_head_dim=8, block_out_channels=(320, 640, 1280, 1280),
    cross_attention_dim=768,
    down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
    up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
    downsample_padding=1, layers_per_block=2, mid_block_scale_factor=1, norm_eps=1e-5, norm_num_groups=32,
    sample_size=64, use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False,
    motion_module_decoder_only=False, motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                              attention_block_types=("Temporal_Self", "Temporal_Self"),
                              temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                              temporal_attention_dim_div=1),
)

VAE_PARAMS = dict(
    embed_dim=4,
    ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4, 4),
                  num_res_blocks=2, attn_resolutions=(), dropout=0.0),
    lossconfig=dict(target="torch.nn.Identity"),
)

SCALE_FACTOR = 0.18215  # diffusion.py:247-249 / instruct_v2v_inference.yaml:15
