"""Process-wide switches of the host side."""

# Route deformable convolutions of the model's hot shape (64->64, 3x3, s=p=d=1, groups=1) to the tcgen05
# implicit-GEMM kernel (bf16 operands, fp32 accumulate).  False: every call takes the fp32 catch-all kernel.
tensor_core = True
