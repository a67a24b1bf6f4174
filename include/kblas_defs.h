/*
 * kblas_defs.h -- constants of the KBLAS batch API, B200-native drop-in.
 *
 * Values are part of the binary contract and therefore identical to the
 * reference (ecrc/kblas-gpu include/kblas_defs.h:25-53): KBLAS_Success is 1,
 * KBLAS_UnknownError is 0 and every other error code is negative.
 */
#ifndef KBLAS_B200_DEFS_H
#define KBLAS_B200_DEFS_H

/* character selectors (reference include/kblas_defs.h:25-34) */
#define KBLAS_Lower   'L'
#define KBLAS_Upper   'U'
#define KBLAS_Left    'L'
#define KBLAS_Right   'R'
#define KBLAS_Trans   'T'
#define KBLAS_NoTrans 'N'
#define KBLAS_Unit    'U'
#define KBLAS_NonUnit 'N'
#define KBLAS_Symm    'S'
#define KBLAS_NonSymm 'N'

/* return codes (reference include/kblas_defs.h:36-49) */
#define KBLAS_Success                1
#define KBLAS_UnknownError           0
#define KBLAS_NotSupported          -1
#define KBLAS_NotImplemented        -2
#define KBLAS_cuBLAS_Error          -3
#define KBLAS_WrongConfig           -4
#define KBLAS_CUDA_Error            -5
#define KBLAS_InsufficientWorkspace -6
#define KBLAS_Error_Allocation      -7
#define KBLAS_Error_Deallocation    -8
#define KBLAS_Error_NotInitialized  -9
#define KBLAS_Error_WrongInput     -10
#define KBLAS_MAGMA_Error          -11
#define KBLAS_SVD_NoConvergence    -12

/* limits (reference include/kblas_defs.h:53-55) */
#define MAX_NGPUS      (16)
#define MAX_STREAMS    (1)
#define KBLAS_NSTREAMS 10

#endif /* KBLAS_B200_DEFS_H */
