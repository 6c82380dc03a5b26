// dlaf/init.h — dlaf::initialize / finalize / ScopedInitializer (reference: include/dlaf/init.h:57-110,
// src/init.cpp:366-429). There is no pika runtime to start: initialization selects the CUDA device.
#pragma once

#include <dlaf_c/init.h>

namespace dlaf {

inline void initialize(int argc, const char* const argv[]) {
  dlaf_initialize(0, nullptr, argc, const_cast<const char**>(argv));
}
inline void initialize() {
  dlaf_initialize(0, nullptr, 0, nullptr);
}
inline void finalize() {
  dlaf_finalize();
}

struct [[nodiscard]] ScopedInitializer {
  ScopedInitializer(int argc, const char* const argv[]) { initialize(argc, argv); }
  ScopedInitializer() { initialize(); }
  ~ScopedInitializer() { finalize(); }
  ScopedInitializer(const ScopedInitializer&) = delete;
  ScopedInitializer& operator=(const ScopedInitializer&) = delete;
};

}  // namespace dlaf
