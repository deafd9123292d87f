"""Process-wide switches of the host side."""

# Route deformable convolutions of the model's hot shape (64->64, 3x3, s=p=d=1, groups=1) to the tcgen05
# implicit-GEMM kernel (bf16 operands, fp32 accumulate).  False: every call takes the fp32 catch-all kernel.
tensor_core = True

# Gather of the fused alignment path (MVDualAttAlignment: head fields -> DCN):
#   "tex"   bilinear footprint fetched by the texture units (cdfo_dcn_tex_sm100_fwd; 8-bit filter weights, fp16 operands)
#   "exact" fp32 coordinate / weight arithmetic of the reference with LDG gathers (cdfo_dcn_sm100_fwd; bf16 operands)
dcn_gather = "tex"
