// TEMPORARY: entry points declared in include/mhb200.h whose kernels are not written yet.
#include "common.cuh"
#define MHB_STUB(name, ...) extern "C" int32_t name(__VA_ARGS__) { mhb::set_error(#name ": not built yet"); return MHB_E_UNSUPPORTED; }
MHB_STUB(mhb_window_order_f32, const float*, const mhb_windows*, const int32_t*, const double*, int32_t, const mhb_table*, void*)
MHB_STUB(mhb_window_order_f64, const double*, const mhb_windows*, const int32_t*, const double*, int32_t, const mhb_table*, void*)
MHB_STUB(mhb_window_spectral_f32, const float*, const mhb_windows*, double, const int32_t*, const double*, int32_t, const mhb_table*, void*)
MHB_STUB(mhb_fft_c128, const double*, int32_t, int64_t, int32_t, int32_t, double*, void*)
MHB_STUB(mhb_window_psd_f32, const float*, const mhb_windows*, void*, int32_t, void*)
MHB_STUB(mhb_haversine_elementwise, const double*, const double*, const double*, const double*, int64_t, double*, void*)
MHB_STUB(mhb_haversine_vector, double, double, const double*, const double*, int64_t, double*, void*)
MHB_STUB(mhb_haversine_outer, const double*, const double*, int64_t, const double*, const double*, int64_t, double*, void*)
MHB_STUB(mhb_successive_distance, const double*, const double*, const int64_t*, int64_t, int64_t, double*, void*)
MHB_STUB(mhb_location_segments, const double*, const double*, const int64_t*, const int64_t*, int64_t, const double*, double, double, int64_t, double*, int64_t*, void*)
MHB_STUB(mhb_label_stats, const int64_t*, int64_t, int64_t, int64_t, int64_t, int64_t*, double*, void*)
MHB_STUB(mhb_minmax_i64, const int64_t*, int64_t, int64_t*, void*)
