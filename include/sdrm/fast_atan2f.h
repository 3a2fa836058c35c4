/* Drop-in for reference src/math/fast_atan2f.h:4 (host evaluation of the same table arctangent the kernels use). */
#ifndef SDRM_FAST_ATAN2F_H
#define SDRM_FAST_ATAN2F_H

float fast_atan2f(float y, float x);

#endif
