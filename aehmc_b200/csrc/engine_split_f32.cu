#include "engine_split.inl"

namespace b2h {
template int run_split_g<float>(b2h_ctx*, EngineView<float>&, const EnginePlan&, const b2h_model*, const b2h_metric*,
                              const b2h_cfg*, i64, int, void*, i64, int*, int, bool);
}
