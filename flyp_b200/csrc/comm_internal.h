// comm_internal.h — what api.cu needs from comm.cu beyond the public C ABI (include/flyp_clip.h).
#pragma once
#include "../../include/flyp_clip.h"
#include "peer.cuh"

namespace flyp {

// Extra work the pack kernel of a gather does for the forward that follows it (one pass over the local rows serves the
// exchange slots, their fp16 copies AND the positive logits): t2[r] = scale log2(e) <img_r, txt_r> and pos[r] =
// row_offset + r for r < n_rows, padding rows [n_rows, n_pad) get (-inf, -1); zero_words[0 .. n_zero) are cleared.
struct PackExtra {
    const float* scale;
    float* t2;
    int* pos;
    int n_pad;
    int row_offset;
    int* zero_words;
    int n_zero;
};

int comm_gather(flyp_comm* c, const void* img, const void* txt, int n_rows, int dim, int dtype, const PackExtra* extra,
                flyp_gathered_t* out, void* stream);
// Destinations of this rank's d(logit_scale) partial of step `seq` (published by the last CTA of the first sweep).
int comm_scalar_push_target(flyp_comm* c, uint32_t seq, PeerPush* out);

// Reduce-scatter of the text gradient (kept-dS backward): where this rank's fp32 partials of rank q's rows go
// (out_rank[world], local or peer-mapped), the release of the step's flag behind the kernel that wrote them, and the
// sum of the W slots of the own buffer, times mul, into out[n_rows, dim] (fp32 or bf16) once every rank's flag has arrived.
int comm_rs_targets(flyp_comm* c, uint32_t seq, int n_rows, int dim, float** out_rank);
int comm_rs_signal(flyp_comm* c, uint32_t seq, void* stream);
int comm_rs_reduce(flyp_comm* c, uint32_t seq, int n_rows, int dim, void* out, int out_fp32, float mul, void* stream);
int comm_world(const flyp_comm* c);
int comm_rs_min_rows(const flyp_comm* c);   // rows per rank from which the kept-dS product + reduce-scatter is used

void set_error(int code, const char* fmt, ...);   // api.cu: thread-local message returned by flyp_last_error()

}  // namespace flyp
