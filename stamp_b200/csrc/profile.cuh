// Optional per-category CUDA-event timing of the library's own launches (off by default).
// bench.py switches it on for one pass to measure the dominant kernel's average duration live,
// on the stream the kernels are launched on.
#pragma once
#include <cuda_runtime.h>

namespace sb {

enum ProfCat : int { PROF_GEMM = 0, PROF_ATTN = 1, PROF_ROWOP = 2, PROF_MACENKO = 3, PROF_POOL = 4, PROF_NCAT = 5 };

bool prof_enabled();
// record an event pair around a launch: call begin before and end after the <<<>>> statement
void prof_begin(int cat, cudaStream_t stream);
void prof_end(int cat, double work, cudaStream_t stream);  // work = algorithmic FLOPs or bytes

struct ProfScope {
    int cat; double work; cudaStream_t s; bool on;
    ProfScope(int c, double w, cudaStream_t st) : cat(c), work(w), s(st), on(prof_enabled()) { if (on) prof_begin(cat, s); }
    ~ProfScope() { if (on) prof_end(cat, work, s); }
};

}  // namespace sb
