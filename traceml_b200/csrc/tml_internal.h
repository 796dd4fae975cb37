// Private glue between the translation units of libtraceml_b200.so (not part of the ABI).
#pragma once

#include "../../include/traceml_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

// sets tml_last_error() for the calling thread and returns `code`
int tml_set_error_(int code, const char* fmt, ...);
// storage for tml_reduce_run's workspace inside the context (owned by tml_summary.cpp)
void** tml_run_ws_slot_(tml_ctx* ctx);
void tml_run_ws_free_(void* ws);

#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
// ---------------------------------------------------------------------------------------------
// JSON text is built from many short-lived strings (rule engines, section objects: ~10^3 per
// summary).  They all die when the extern "C" entry point returns, so they come from a per-thread
// bump arena that is rewound when the outermost entry point leaves: allocation is a pointer
// increment, deallocation nothing.
#include <cstddef>
#include <cstdlib>
#include <new>
#include <string>

namespace tml_json {

struct Arena {
  static constexpr size_t BLOCK = 1u << 16;
  static constexpr int MAX_BLOCKS = 256;
  char* blocks[MAX_BLOCKS];
  int nblocks = 0, cur = 0, depth = 0;
  size_t off = 0;
  void* big[MAX_BLOCKS];  // requests that do not fit a block
  int nbig = 0;
  void* take(size_t n) {
    n = (n + 15u) & ~(size_t)15u;
    if (n > BLOCK / 4 || depth == 0) {  // oversized, or outside any scope (static initialisers)
      void* p = std::malloc(n);
      if (!p) throw std::bad_alloc();
      if (depth > 0 && nbig < MAX_BLOCKS) big[nbig++] = p;  // else: leaked on purpose (never in practice)
      return p;
    }
    if (nblocks == 0 || off + n > BLOCK) {
      if (nblocks > 0 && cur + 1 < nblocks) { ++cur; }
      else {
        if (nblocks == MAX_BLOCKS) throw std::bad_alloc();
        char* b = (char*)std::malloc(BLOCK);
        if (!b) throw std::bad_alloc();
        blocks[nblocks] = b; cur = nblocks++;
      }
      off = 0;
    }
    void* p = blocks[cur] + off;
    off += n;
    return p;
  }
  void rewind() {
    cur = 0; off = 0;
    for (int i = 0; i < nbig; ++i) std::free(big[i]);
    nbig = 0;
  }
  ~Arena() { rewind(); for (int i = 0; i < nblocks; ++i) std::free(blocks[i]); }
};

inline Arena& arena() {
  static thread_local Arena a;
  return a;
}

struct Scope {  // first statement of every entry point that builds JSON
  Arena& a;
  Scope() : a(arena()) { ++a.depth; }
  ~Scope() { if (--a.depth == 0) a.rewind(); }
  Scope(const Scope&) = delete;
  Scope& operator=(const Scope&) = delete;
};

template <class T>
struct Alloc {
  typedef T value_type;
  Alloc() noexcept {}
  template <class U> Alloc(const Alloc<U>&) noexcept {}
  T* allocate(size_t n) { return (T*)arena().take(n * sizeof(T)); }
  void deallocate(T*, size_t) noexcept {}
  template <class U> bool operator==(const Alloc<U>&) const noexcept { return true; }
  template <class U> bool operator!=(const Alloc<U>&) const noexcept { return false; }
};

typedef std::basic_string<char, std::char_traits<char>, Alloc<char>> Str;

}  // namespace tml_json
#endif
