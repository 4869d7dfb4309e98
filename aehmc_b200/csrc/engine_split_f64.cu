#include "engine_split.inl"

namespace b2h {
template int run_split_g<double>(b2h_ctx*, EngineView<double>&, const EnginePlan&, const b2h_model*, const b2h_metric*,
                              const b2h_cfg*, i64, int, void*, i64, int*, int, bool);
}
