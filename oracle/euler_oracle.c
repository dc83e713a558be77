/* TEST INFRASTRUCTURE ONLY -- builds the float and double instantiations of the CPU restatement in euler_impl.inc.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may load the resulting library. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL double
#define SFX _f64
#define MLOG log
#define MSQRT sqrt
#define MABS fabs
#define MCBRT cbrt
#define MEXP exp
#define MPOW pow
#define MACOS acos
#define MASIN asin
#define MCOS cos
#include "euler_impl.inc"
#undef REAL
#undef SFX
#undef MLOG
#undef MSQRT
#undef MABS
#undef MCBRT
#undef MEXP
#undef MPOW
#undef MACOS
#undef MASIN
#undef MCOS

#define REAL float
#define SFX _f32
#define MLOG logf
#define MSQRT sqrtf
#define MABS fabsf
#define MCBRT cbrtf
#define MEXP expf
#define MPOW powf
#define MACOS acosf
#define MASIN asinf
#define MCOS cosf
#include "euler_impl.inc"
