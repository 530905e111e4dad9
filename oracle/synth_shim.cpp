// Test infrastructure: the synthetic failures.jsonl-shaped row generator (kakveda_b200/csrc/synth.cpp, plain host C++)
// built WITHOUT the CUDA library, so that `bench.py --impl reference` can create its inputs without loading
// libkakveda_b200.so (the reference arm must not touch the product library).  Built by __graft_entry__.build() into
// oracle/_build/libkvsynth.so with g++.
#include "../kakveda_b200/csrc/kv_internal.h"

int kv_fail(int code, const char *, ...) { return code; }
void kv_clear_error() {}

#include "../kakveda_b200/csrc/synth.cpp"
