// Config.h -- host-side mirror of cimbar::Config (reference: src/lib/cimb_translator/Config.h:7-175, GridConf.h:8-190).
// Same static accessor names and meaning; the numbers come from the C ABI (cb200_mode_info), so there is one mode table.
// Like the reference, the active mode is thread_local: a worker thread calls Config::update(mode) itself.
#pragma once
#include "../../include/cb200.h"
#include <stdexcept>
#include <string>

namespace cb200 {
namespace cimbar {

class Config
{
protected:
	static cb200_info& active_conf()
	{
		static thread_local cb200_info cc = temp_conf(68);
		return cc;
	}

public:
	static cb200_info temp_conf(int mode_val = 0)
	{
		cb200_info info;
		if (mode_val == 0) mode_val = 68;   // Config.h:39-42: default -> Conf8x8
		if (cb200_mode_info(mode_val, &info) != CB200_OK)
			cb200_mode_info(68, &info);
		return info;
	}

	static void update(int mode_val = 0) { active_conf() = temp_conf(mode_val); }

	static int mode_val() { return active_conf().mode_val; }
	static bool dark() { return true; }
	static bool legacy_mode() { return active_conf().legacy_mode != 0; }
	static unsigned color_mode() { return active_conf().legacy_mode ? 0 : 1; }
	static unsigned color_bits() { return active_conf().color_bits; }
	static unsigned symbol_bits() { return active_conf().symbol_bits; }
	static unsigned bits_per_cell() { return active_conf().color_bits + active_conf().symbol_bits; }
	static unsigned ecc_bytes() { return active_conf().ecc_bytes; }
	static unsigned ecc_block_size() { return active_conf().ecc_block_size; }
	static unsigned image_size_x() { return active_conf().image_size_x; }
	static unsigned image_size_y() { return active_conf().image_size_y; }
	static unsigned anchor_size() { return 30; }
	static constexpr unsigned cell_size() { return 8; }
	static unsigned total_cells() { return active_conf().total_cells; }
	static unsigned capacity(unsigned bitspercell = 0)
	{
		if (!bitspercell) bitspercell = bits_per_cell();
		return total_cells() * bitspercell / 8;
	}
	static unsigned interleave_blocks() { return ecc_block_size(); }
	static unsigned interleave_partitions() { return 2; }
	static unsigned fountain_chunks_per_frame(unsigned = 0) { return active_conf().chunks_per_frame; }
	static unsigned fountain_chunk_size(unsigned = 0) { return active_conf().chunk_size; }
	static unsigned compression_level() { return 16; }
};

}  // namespace cimbar
}  // namespace cb200
