// Compile-time constants of the model / kernels, shared by the CUDA sources and the CPU warp emulation (tools/emu).
#pragma once
#define HID 32          // hidden width of message / pass MLPs
#define HD 48           // h_dim
#define ED 48           // e_dim
#define EDR 16          // numerical rank of the radial-descriptor family at FP32 precision (see epnn_create: rbf_basis)
#define UPD_IN 80       // [h | M]
#define SMALL_MAX 48    // systems with n <= SMALL_MAX are packed into warp-private "bundles" (epnn_bundle.cu)
#define BUNDLE_ATOMS SMALL_MAX   // max atoms of one bundle (whole systems only)
#define MAX_SPECIES 16
#define CELL_MIN 512     // systems with more atoms than this build their neighbour list through a cell list
