#include "engine_fused.inl"

namespace b2h {
template int launch_fused_g<float>(cudaStream_t, const EngineView<float>&, const b2h_model*, i64, int, bool);
}
