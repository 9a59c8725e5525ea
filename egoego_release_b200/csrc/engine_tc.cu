// Placeholder until the tcgen05 engine lands (replaced in the next milestone).
#include "engine_tc.cuh"
namespace egoego {
struct TcImpl {};
TcEngine::TcEngine() : impl_(nullptr) {}
TcEngine::~TcEngine() {}
int TcEngine::init(const TcWeights&, cudaStream_t) { set_error("tcgen05 engine not built yet"); return 1; }
int TcEngine::clear_staging(int, cudaStream_t) { return 1; }
int TcEngine::stage(const float*, int, int, bool, int, int, cudaStream_t, int64_t*) { return 1; }
int TcEngine::prepare_cond(int, int, cudaStream_t, int64_t*) { return 1; }
int TcEngine::denoiser(int, int, TSrc, const float*, float*, cudaStream_t, int64_t*) { return 1; }
void TcEngine::stage_targets(__nv_bfloat16**, __nv_bfloat16**, int*) {}
int TcEngine::launches_per_denoiser() const { return 0; }
}
