// oracle/shim/tiny_obj_loader.h -- TEST INFRASTRUCTURE ONLY.
//
// The reference's mesh.cpp (srcCommon/scene/geometry/mesh.cpp:3,192) includes <tiny_obj_loader.h> from the SYSTEM and
// calls the six-argument tinyobj::LoadObj(attrib, shapes, materials, &warn, &err, path) of tinyobjloader >= 1.3, a
// dependency that is neither vendored by the reference nor pinned by its build (CMakeLists.txt:46 find_package).
// The only tinyobjloader under /root/reference is v1.2.0 (srcVulkan/dep/slang/external/tinyobjloader, a nested
// submodule of Slang), whose LoadObj has one message string.  This shim lets the UNMODIFIED mesh.cpp compile against
// that copy: it includes the vendored header from where it lies and adds the newer overload, which forwards to the
// 1.2.0 parser and sorts its single message string into `warn` (the load succeeded) or `err` (it failed), which is
// how the later versions split it.  Parity of OBJ ingestion is therefore pinned against
// "mesh.cpp + tinyobjloader 1.2.0", the version in the reference's own tree; see DESIGN.md 4.10.
#ifndef RTR_ORACLE_SHIM_TINY_OBJ_LOADER_H
#define RTR_ORACLE_SHIM_TINY_OBJ_LOADER_H

#include RTR_VENDORED_TINYOBJ  // -DRTR_VENDORED_TINYOBJ='"<path>/tiny_obj_loader.h"' (oracle/Makefile)

namespace tinyobj {
inline bool LoadObj(attrib_t* attrib, std::vector<shape_t>* shapes, std::vector<material_t>* materials,
                    std::string* warn, std::string* err, const char* filename) {
    std::string msg;
    const bool ok = LoadObj(attrib, shapes, materials, &msg, filename, static_cast<const char*>(0), true);
    if (ok) { if (warn) *warn = msg; } else { if (err) *err = msg.empty() ? std::string("LoadObj failed\n") : msg; }
    return ok;
}
}  // namespace tinyobj
#endif
