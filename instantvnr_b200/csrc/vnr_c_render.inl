// vnr_c_render.inl -- renderer entry points
VNR_EXPORT int vnr_renderer_create(vnr_volume_t*, vnr_renderer_t**) VNR_TODO("vnr_renderer_create")
VNR_EXPORT void vnr_renderer_release(vnr_renderer_t*) {}
VNR_EXPORT int vnr_renderer_set_size(vnr_renderer_t*, int, int) VNR_TODO("vnr_renderer_set_size")
VNR_EXPORT int vnr_renderer_set_camera(vnr_renderer_t*, const float*, const float*, const float*, float) VNR_TODO("vnr_renderer_set_camera")
VNR_EXPORT int vnr_renderer_set_mode(vnr_renderer_t*, int) VNR_TODO("vnr_renderer_set_mode")
VNR_EXPORT int vnr_renderer_set_sampling_rate(vnr_renderer_t*, float) VNR_TODO("vnr_renderer_set_sampling_rate")
VNR_EXPORT int vnr_renderer_set_density_scale(vnr_renderer_t*, float) VNR_TODO("vnr_renderer_set_density_scale")
VNR_EXPORT int vnr_renderer_reset_accumulation(vnr_renderer_t*) VNR_TODO("vnr_renderer_reset_accumulation")
VNR_EXPORT int vnr_renderer_set_partition(vnr_renderer_t*, int, int) VNR_TODO("vnr_renderer_set_partition")
VNR_EXPORT int vnr_renderer_set_jitter_mode(vnr_renderer_t*, int) VNR_TODO("vnr_renderer_set_jitter_mode")
VNR_EXPORT int vnr_render(vnr_renderer_t*) VNR_TODO("vnr_render")
VNR_EXPORT const float* vnr_map_frame(vnr_renderer_t*) { g_last_error = "vnr_map_frame: not implemented yet"; return nullptr; }
VNR_EXPORT int vnr_renderer_device_frame(vnr_renderer_t*, void**, void*) VNR_TODO("vnr_renderer_device_frame")
VNR_EXPORT int vnr_renderer_stats(vnr_renderer_t*, uint64_t*) VNR_TODO("vnr_renderer_stats")
