// tests/emu/simt_emu.cpp — TEST INFRASTRUCTURE (see simt_emu.hpp).
#include "simt_emu.hpp"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

// ---- context switch (x86-64 SysV): saves the callee-saved registers on the current stack, swaps stack pointers ----
extern "C" void mapad_emu_ctx_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl mapad_emu_ctx_switch
.type mapad_emu_ctx_switch,@function
mapad_emu_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size mapad_emu_ctx_switch,.-mapad_emu_ctx_switch
)");

namespace {

struct GroupSync {
  uint32_t slot[32];
  int arrived = 0;
  uint32_t gen = 0;
};

struct LaneCtx {
  void* site = nullptr;   // return address of the collective the lane is waiting in (stall diagnostics)
  void* sp = nullptr;
  char* stack = nullptr;
  bool done = false;
  int group = 0, lane = 0;
};

struct Sched {
  std::vector<LaneCtx> lanes;
  std::vector<GroupSync> groups;
  const std::function<void(int, int)>* fn = nullptr;
  void* main_sp = nullptr;
  int current = -1;
  int group_size = 1;
  unsigned long long yields = 0;  // explicit back-off yields (a lane polling for a resource is not a divergence stall)
};

thread_local Sched* g_sched = nullptr;

void yield_to_main() {
  Sched* s = g_sched;
  LaneCtx& me = s->lanes[s->current];
  mapad_emu_ctx_switch(&me.sp, s->main_sp);
}

extern "C" void mapad_emu_lane_entry() {
  Sched* s = g_sched;
  LaneCtx& me = s->lanes[s->current];
  (*s->fn)(me.group, me.lane);
  me.done = true;
  yield_to_main();
  abort();  // a finished lane is never resumed
}

void barrier(GroupSync& g, int n, void* site) {
  g_sched->lanes[g_sched->current].site = site;
  const uint32_t gen = g.gen;
  if (++g.arrived == n) {
    g.arrived = 0;
    g.gen = gen + 1;
    return;
  }
  while (g.gen == gen) yield_to_main();
}

constexpr size_t STACK_BYTES = 512 * 1024;

}  // namespace

namespace simt_emu {

void run(int n_groups, int group_size, const std::function<void(int, int)>& fn) {
  Sched s;
  s.fn = &fn;
  s.group_size = group_size;
  s.groups.resize(n_groups);
  s.lanes.resize((size_t)n_groups * group_size);
  for (int g = 0; g < n_groups; ++g)
    for (int l = 0; l < group_size; ++l) {
      LaneCtx& c = s.lanes[(size_t)g * group_size + l];
      c.group = g; c.lane = l;
      c.stack = (char*)aligned_alloc(64, STACK_BYTES);
      // initial frame: six callee-saved registers (zero) + the return address of the first switch
      uintptr_t top = ((uintptr_t)c.stack + STACK_BYTES) & ~(uintptr_t)63;
      void** sp = (void**)top;
      *--sp = nullptr;                              // keeps (rsp + 8) % 16 == 0 at the entry of the lane function
      *--sp = (void*)&mapad_emu_lane_entry;         // `ret` target
      for (int i = 0; i < 6; ++i) *--sp = nullptr;  // r15 r14 r13 r12 rbx rbp
      c.sp = sp;
    }
  Sched* prev = g_sched;
  g_sched = &s;
  size_t remaining = s.lanes.size();
  unsigned long long idle_rounds = 0;
  const char* order = getenv("MAPAD_SIMT_EMU_ORDER");
  const bool reverse = order && !strcmp(order, "reverse");
  while (remaining) {
    uint32_t gens = 0;
    for (auto& g : s.groups) gens += g.gen;
    const size_t before = remaining;
    const unsigned long long yields_before = s.yields;
    for (size_t k = 0; k < s.lanes.size(); ++k) {
      const size_t i = reverse ? s.lanes.size() - 1 - k : k;
      LaneCtx& c = s.lanes[i];
      if (c.done) continue;
      s.current = (int)i;
      mapad_emu_ctx_switch(&s.main_sp, c.sp);
      if (c.done) remaining -= 1;
    }
    uint32_t gens2 = 0;
    for (auto& g : s.groups) gens2 += g.gen;
    idle_rounds = (gens2 == gens && remaining == before && s.yields == yields_before) ? idle_rounds + 1 : 0;
    if (idle_rounds > 1000) {  // no barrier completed and no lane finished: the lanes of a group have diverged
      fprintf(stderr, "simt_emu: stall (divergent collectives)\n");
      for (auto& c : s.lanes) {
        Dl_info di;
        memset(&di, 0, sizeof di);
        if (c.site) dladdr(c.site, &di);
        fprintf(stderr, "  group %d lane %d done=%d last collective at %s+0x%lx\n", c.group, c.lane, (int)c.done, di.dli_fname ? di.dli_fname : "?",
                (unsigned long)((char*)c.site - (char*)di.dli_fbase));
      }
      abort();
    }
  }
  g_sched = prev;
  for (auto& c : s.lanes) free(c.stack);
}

}  // namespace simt_emu

extern "C" {

uint32_t mapad_simt_emu_shfl(uint32_t v, int src, int group_size) {
  Sched* s = g_sched;
  LaneCtx& me = s->lanes[s->current];
  GroupSync& g = s->groups[me.group];
  g.slot[me.lane] = v;
  barrier(g, group_size, __builtin_return_address(0));
  const uint32_t r = g.slot[src & (group_size - 1)];
  barrier(g, group_size, __builtin_return_address(0));
  return r;
}

uint32_t mapad_simt_emu_ballot(int pred, int group_size) {
  Sched* s = g_sched;
  LaneCtx& me = s->lanes[s->current];
  GroupSync& g = s->groups[me.group];
  g.slot[me.lane] = pred ? 1u : 0u;
  barrier(g, group_size, __builtin_return_address(0));
  uint32_t r = 0;
  for (int l = 0; l < group_size; ++l) r |= g.slot[l] << l;
  barrier(g, group_size, __builtin_return_address(0));
  return r;
}

void mapad_simt_emu_yield(void) {
  if (g_sched) { g_sched->yields += 1; yield_to_main(); }
}

void mapad_simt_emu_sync(int group_size) {
  Sched* s = g_sched;
  LaneCtx& me = s->lanes[s->current];
  barrier(s->groups[me.group], group_size, __builtin_return_address(0));
}

}  // extern "C"
