// onef.h -- device generation of 1/f^alpha noise buffers (onef.cu).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "../../include/cmtj_b200.h"

namespace cmtjb {

// one family of series: component `axis` of layer `layer`, one independent series per instance
struct OnefSlot {
  int layer, axis;
  double alpha, std;
};

// The series a junction model needs (VectorAlphaNoise, noise.hpp:259-288: three components for Axis::all,
// one otherwise).  slot_of_layer[l] = first slot of layer l or -1.  Returns the slot count, -1 on an invalid axis.
int onef_slots_of(const cmtjb_layer_t *layers, int n_layers, OnefSlot *slots, int *slot_of_layer);

// out[(s * n + t) * Bc + g] = BufferedAlphaNoise result[t] (raw float32: neither scaled nor normalised) of slot s
// for global element elem_global0 + g.  Stream-ordered; scratch is allocated and freed on the stream.
int onef_generate(int device, cudaStream_t stream, long long n, const OnefSlot *slots, int n_slots, unsigned long long seed,
                  unsigned long long elem_global0, long long Bc, float *out);

// bytes of FFT scratch a generation of n_series series of n samples holds at a time
long long onef_group_size(long long n, long long n_series);
size_t onef_scratch_bytes(long long n, long long group);

void kasdin_weights(long long n, double alpha, std::vector<float> &h);

}  // namespace cmtjb
