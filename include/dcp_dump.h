/* dcp_dump.h -- on-disk container for cross-validation against a real deal.II build.
 *
 * The reference cannot be built in the image this library was developed in, so parity with it is pinned only
 * indirectly.  A user who HAS a deal.II/Trilinos build of 3D-DyCorePlanet can close that gap: dump
 * the inputs of the hot path (dof maps, constraint lines, sparsity patterns, mapping records, reference tables,
 * solution vectors) and the matrices / right-hand sides the reference assembled from them into one file with the
 * few calls below (header-only, plain C, no dependency), and run
 *     DCP_REFERENCE_DUMP=file.dcpd python -m pytest tests/test_external_reference.py -m gpu
 * which feeds the same inputs to the CUDA path and compares at the tolerances of BASELINE.json.
 * INTEGRATION.md section 5 lists the array names and shows the deal.II-side loop.
 *
 * Layout (little endian): 8 bytes magic "DCPDUMP1", then records
 *     int32 name_len | name bytes | int32 dtype (0 f64, 1 i32, 2 i64, 3 i8, 4 i16) | int64 count | payload | pad to 8 bytes
 * until end of file.  Scalars are 1-element i64 arrays named "scalar:<name>"; the spec string is an i8 array "spec".
 */
#ifndef DCP_DUMP_H
#define DCP_DUMP_H
#include <stdint.h>
#include <stdio.h>
#include <string.h>

enum { DCP_DUMP_F64 = 0, DCP_DUMP_I32 = 1, DCP_DUMP_I64 = 2, DCP_DUMP_I8 = 3, DCP_DUMP_I16 = 4 };

static inline FILE* dcp_dump_open(const char* path) {
  FILE* f = fopen(path, "wb");
  if (f && fwrite("DCPDUMP1", 1, 8, f) != 8) {
    fclose(f);
    return NULL;
  }
  return f;
}

/* returns 0 on success */
static inline int dcp_dump_array(FILE* f, const char* name, int32_t dtype, int64_t count, const void* data) {
  static const int elem[5] = {8, 4, 8, 1, 2};
  static const char zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (!f || !name || dtype < 0 || dtype > 4 || count < 0 || (count > 0 && !data)) return 1;
  const int32_t len = (int32_t)strlen(name);
  const size_t bytes = (size_t)count * (size_t)elem[dtype];
  const size_t head = 4 + (size_t)len + 4 + 8;
  const size_t pad = (8 - (head + bytes) % 8) % 8;
  if (fwrite(&len, 4, 1, f) != 1 || fwrite(name, 1, (size_t)len, f) != (size_t)len || fwrite(&dtype, 4, 1, f) != 1 ||
      fwrite(&count, 8, 1, f) != 1)
    return 1;
  if (bytes && fwrite(data, 1, bytes, f) != bytes) return 1;
  if (pad && fwrite(zeros, 1, pad, f) != pad) return 1;
  return 0;
}

static inline int dcp_dump_scalar(FILE* f, const char* name, int64_t value) {
  char full[256];
  snprintf(full, sizeof full, "scalar:%s", name);
  return dcp_dump_array(f, full, DCP_DUMP_I64, 1, &value);
}

static inline int dcp_dump_close(FILE* f) { return f ? fclose(f) : 1; }
#endif
