// role-dependent constants of the persistent tcgen05 GEMM (derived from P2_PROD_WARPS; see x3d_pw_tc2.cu)
#define P2_PROD_PASSES (TC_BM / (P2_PROD_WARPS * 4))   /* rows per thread and chunk: 128 rows / (warps x 4 rows per warp) */
#define P2_EPI_WARPS 8
#define P2_PROD_THREADS (P2_PROD_WARPS * 32)
#define P2_EPI_THREADS (P2_EPI_WARPS * 32)
#define P2_MMA_WARP P2_PROD_WARPS
#define P2_EPI_WARP0 (P2_PROD_WARPS + 4)          /* roles are warpgroup (4-warp) aligned for setmaxnreg */
#define P2_THREADS ((P2_PROD_WARPS + 4 + P2_EPI_WARPS) * 32)
