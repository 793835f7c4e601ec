// TEST INFRASTRUCTURE: stand-in for graphlily/synthesizer/overlay_synthesizer.h
// (/root/reference/graphlily/synthesizer/overlay_synthesizer.h: writes a Vitis project and runs the
// bitstream build).  On B200 the kernels are built by nvcc into libgraphlily_b200.so, so "synthesize"
// has nothing to do; the class exists so that the reference's test files compile unmodified.
#ifndef GLB_REF_COMPAT_OVERLAY_SYNTHESIZER_H_
#define GLB_REF_COMPAT_OVERLAY_SYNTHESIZER_H_
#include <cstdint>
#include <string>

namespace graphlily {
const std::string proj_folder_name = "glb_b200_proj";   // the tests `rm -rf` it afterwards
namespace synthesizer {
class OverlaySynthesizer {
public:
    OverlaySynthesizer(uint32_t, uint32_t, uint32_t, uint32_t) {}
    void set_target(const std::string &) {}
    void synthesize() {}
};
}  // namespace synthesizer
}  // namespace graphlily
#endif
