// detail.h -- per-thread state shared by the Decoder / CimbReader / CimbDecoder mirrors.
//
// The reference keeps two things per THREAD, not per object, and its own entry points rely on that:
//   * the colour correction matrix -- `static thread_local color_correction` behind CimbDecoder::internal_ccm()
//     (src/lib/cimb_translator/CimbDecoder.cpp:69-85): every CimbDecoder of a thread reads and writes the same CCM, so a facade
//     that builds a fresh `Decoder` per frame (cimbar_recv_js.cpp:164) still carries a fitted matrix from frame to frame;
//   * the active mode (thread_local Config, Config.h:11-15).
// The mirrors therefore share one CCM per thread (pushed into the device context before a call, pulled back after it) and one
// cached device context per (thread, device, mode): constructing a mirror object costs nothing after the first one.
#pragma once
#include "../../include/cb200.h"

#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>

namespace cb200 {
namespace detail {

struct ThreadCcm { bool active = false; float m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; unsigned version = 0; /* bumped on every change */ };
inline ThreadCcm& thread_ccm() { static thread_local ThreadCcm c; return c; }

struct ContextCache
{
	std::map<std::tuple<int, int>, std::pair<cb200_ctx*, int>> entries;    // (device, mode_val) -> (context, max_frames)
	~ContextCache() { for (auto& kv : entries) cb200_destroy(kv.second.first); }
};
inline ContextCache& context_cache() { static thread_local ContextCache c; return c; }

// the calling thread's context for (device, mode_val), able to take at least `frames` frames per call
inline cb200_ctx* thread_context(int device, int mode_val, int frames = 1)
{
	auto& cache = context_cache().entries;
	auto key = std::make_tuple(device, mode_val);
	auto it = cache.find(key);
	if (it != cache.end() and it->second.second >= frames) return it->second.first;
	if (it != cache.end()) { cb200_destroy(it->second.first); cache.erase(it); }
	cb200_ctx* ctx = nullptr;
	if (cb200_create(&ctx, device, mode_val, frames) != CB200_OK)
		throw std::runtime_error(std::string("cb200_create: ") + cb200_last_error());
	cache[key] = {ctx, frames};
	return ctx;
}

inline void push_ccm(cb200_ctx* ctx)
{
	const ThreadCcm& t = thread_ccm();
	if (cb200_set_ccm(ctx, t.active ? t.m : nullptr) != CB200_OK) throw std::runtime_error(std::string("cb200_set_ccm: ") + cb200_last_error());
}
inline void pull_ccm(cb200_ctx* ctx)
{
	ThreadCcm& t = thread_ccm();
	float m[9];
	int rc = cb200_get_ccm(ctx, m);
	if (rc < 0) throw std::runtime_error(std::string("cb200_get_ccm: ") + cb200_last_error());
	const bool changed = (rc == 1) != t.active or (rc == 1 and std::memcmp(t.m, m, sizeof(m)) != 0);
	t.active = rc == 1;
	if (t.active) std::memcpy(t.m, m, sizeof(m));
	if (changed) ++t.version;
}

}  // namespace detail
}  // namespace cb200
