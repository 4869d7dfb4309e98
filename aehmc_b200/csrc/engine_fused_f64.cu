#include "engine_fused.inl"

namespace b2h {
template int launch_fused_g<double>(cudaStream_t, const EngineView<double>&, const b2h_model*, i64, int, bool);
}
