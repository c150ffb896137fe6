/* Minimal stand-in for torch-0.4's <TH/TH.h>, just enough for the reference's CPU RoI-pooling
 * source (lib/layer_utils/roi_pooling/src/roi_pooling.c) to compile UNMODIFIED into oracle/_ref/.
 * Test infrastructure only. */
#ifndef SIS3D_TH_SHIM_H
#define SIS3D_TH_SHIM_H
#include <float.h>
typedef struct THFloatTensor { float *data; long size[8]; } THFloatTensor;
static inline float *THFloatTensor_data(THFloatTensor *t) { return t->data; }
static inline long THFloatTensor_size(THFloatTensor *t, int d) { return t->size[d]; }
#endif
