// C ABI of the stand-in problem builder (see include/dcp_harness.h).
#include "../../../include/dcp_harness.h"

#include <string>

#include "problem.hpp"

struct dcph_problem {
  std::unique_ptr<dcph::Problem> P;
  std::vector<std::string> names, scalar_names;
};

static thread_local std::string g_err;

extern "C" {

dcph_problem* dcph_create(const char* spec) {
  try {
    auto* h = new dcph_problem;
    h->P = dcph::build_problem(dcph::parse_spec(spec ? spec : ""));
    for (auto& kv : h->P->arrays) h->names.push_back(kv.first);
    for (auto& kv : h->P->scalars) h->scalar_names.push_back(kv.first);
    return h;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

void dcph_destroy(dcph_problem* p) { delete p; }

int dcph_array(const dcph_problem* p, const char* name, const void** data, int64_t* count, int* dtype) {
  auto it = p->P->arrays.find(name);
  if (it == p->P->arrays.end()) {
    g_err = std::string("no such array: ") + name;
    return 1;
  }
  *data = it->second.p;
  *count = it->second.n;
  *dtype = it->second.dtype;
  return 0;
}

int64_t dcph_scalar(const dcph_problem* p, const char* name) {
  auto it = p->P->scalars.find(name);
  if (it == p->P->scalars.end()) {
    g_err = std::string("no such scalar: ") + name;
    return -1;
  }
  return it->second;
}

int dcph_n_arrays(const dcph_problem* p) { return (int)p->names.size(); }
const char* dcph_array_name(const dcph_problem* p, int i) { return p->names[i].c_str(); }
int dcph_n_scalars(const dcph_problem* p) { return (int)p->scalar_names.size(); }
const char* dcph_scalar_name(const dcph_problem* p, int i) { return p->scalar_names[i].c_str(); }
const char* dcph_last_error(void) { return g_err.c_str(); }
}
