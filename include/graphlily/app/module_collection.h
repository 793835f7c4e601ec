// graphlily-b200: a set of modules sharing one runtime.
//
// Mirrors /root/reference/graphlily/app/module_collection.h:15-114: owns the modules, and
// set_up_runtime gives every module the same device context (there: one cl::Context with a kernel
// and an out-of-order queue per module; here: one CUDA device + one stream, so launches of
// different modules are ordered without the finish() calls the reference needs).
#ifndef GRAPHLILY_MODULE_COLLECTION_H_
#define GRAPHLILY_MODULE_COLLECTION_H_

#include <cassert>
#include <memory>
#include <string>
#include <vector>

#include "graphlily/global.h"
#include "graphlily/module/base_module.h"

namespace graphlily {
namespace app {

using namespace module;

class ModuleCollection {
protected:
    std::vector<BaseModule *> modules_;
    uint32_t num_modules_ = 0;
    std::vector<std::string> kernel_names_;
    std::string target_ = "hw";
    std::shared_ptr<Runtime> runtime_;

public:
    ModuleCollection() {}
    ModuleCollection(const ModuleCollection &) = delete;
    ModuleCollection &operator=(const ModuleCollection &) = delete;
    virtual ~ModuleCollection() {
        for (size_t i = 0; i < num_modules_; i++) delete modules_[i];
    }

    void add_module(BaseModule *module) {
        modules_.push_back(module);
        kernel_names_.push_back(module->get_kernel_name());
        num_modules_++;
    }

    void set_target(std::string target) {
        assert(target == "sw_emu" || target == "hw_emu" || target == "hw");
        target_ = target;
    }

    // The path named the bitstream in the reference; ignored here.
    void set_up_runtime(std::string /*xclbin_file_path*/) {
        if (!runtime_) runtime_ = Runtime::create_from_env();
        for (size_t i = 0; i < num_modules_; i++) modules_[i]->set_runtime(runtime_);
    }
    void set_runtime(std::shared_ptr<Runtime> runtime) {
        runtime_ = runtime;
        for (size_t i = 0; i < num_modules_; i++) modules_[i]->set_runtime(runtime_);
    }
    void finish() { runtime_->finish(); }
};

}  // namespace app
}  // namespace graphlily

#endif  // GRAPHLILY_MODULE_COLLECTION_H_
