// tools/umma_desc_check.cu — host-only cross-check: the shared-memory and instruction descriptors hand-packed in
// csrc/npw_ozaki_i8.cu (umma_desc_k128, umma_idesc_i8) against the bitfield structs of the vendored CUTLASS headers
// (cute/arch/mma_sm100_desc.hpp).  tests/test_i8emu_protocol.py compiles and runs it when the headers are present.
#include <cstdio>
#include <cstdint>
#include <cute/arch/mma_sm100_desc.hpp>
#include <cute/numeric/numeric_types.hpp>
#include <cute/layout.hpp>
#include <cute/swizzle.hpp>
#include <cute/swizzle_layout.hpp>
static uint64_t my_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
static uint32_t my_idesc(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
int main() {
  using namespace cute;
  UMMA::SmemDescriptor d;
  d.desc_ = 0;
  d.version_ = 1; d.lbo_mode_ = 0; d.layout_type_ = uint8_t(UMMA::LayoutType::SWIZZLE_128B);
  uint32_t addr = 0x12340;
  d.start_address_ = static_cast<uint16_t>(addr >> 4);
  d.base_offset_ = 0;
  d.stride_byte_offset_ = 1024 >> 4;
  d.leading_byte_offset_ = 1;
  printf("cutlass smem desc %016llx\nmine              %016llx\n", (unsigned long long)d.desc_, (unsigned long long)my_desc(addr));
  auto id = UMMA::make_instr_desc<int8_t, int8_t, int32_t, 128, 64, UMMA::Major::K, UMMA::Major::K>();
  printf("cutlass idesc %08x\nmine          %08x\n", (unsigned)uint32_t(id), my_idesc(128, 64));
  // the byte placement tools/tcgen05_i8_probe.cu writes by hand (and TMA's SWIZZLE_128B produces) against CuTe's K-major
  // SW128 atom for 8-bit elements: Swizzle<3,4,3> o (8 rows x 128 bytes, row stride 128)
  auto atom = composition(Swizzle<3, 4, 3>{}, Layout<Shape<_8, _128>, Stride<_128, _1>>{});
  int bad = 0;
  for (int r = 0; r < 8; ++r)
    for (int k = 0; k < 128; ++k)
      if (r * 128 + (((k / 16) ^ (r % 8)) * 16) + k % 16 != int(atom(r, k))) ++bad;
  printf("swizzle formula mismatches: %d\n", bad);
  return (d.desc_ == my_desc(addr) && uint32_t(id) == my_idesc(128, 64) && bad == 0) ? 0 : 1;
}
