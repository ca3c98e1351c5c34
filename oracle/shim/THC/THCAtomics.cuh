/* empty shim: lets oracle/ref_msda_driver.cu include the reference's ms_deform_im2col_cuda.cuh (which only needs CUDA's built-in atomicAdd) without ATen/THC. Test infrastructure. */
