// N2 -- bulk insert of a crown layer into a GeoPackage.
//
// The reference writes its crown layers through geopandas / fiona / OGR (helpers.py:592-599 the stitched
// layer, postprocessing.py:903-936 the processed layer).  A layer of one image holds tens of thousands of
// polygons; building one geometry blob per crown and binding it from the Python interpreter costs more wall
// clock than the whole device side of the image.  This is host code (no kernel): the Python writer creates the
// GeoPackage tables and metadata rows (gpkg.py), then this function appends the features inside one
// transaction -- GeoPackageBinary header + envelope + WKB polygon straight from the (V,2) float64 vertex array
// and the ring offsets.  ctypes releases the GIL for the call, so the writer thread runs beside the decoder
// and the device stage of the next image.  SQLite is reached through dlopen("libsqlite3.so.0") -- the library
// Python's own sqlite3 module is linked against -- so the build needs no headers.
#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

struct Sqlite {
  void* h = nullptr;
  int (*open)(const char*, void**) = nullptr;
  int (*close)(void*) = nullptr;
  int (*exec)(void*, const char*, int (*)(void*, int, char**, char**), void*, char**) = nullptr;
  int (*prepare)(void*, const char*, int, void**, const char**) = nullptr;
  int (*bind_blob)(void*, int, const void*, int, void (*)(void*)) = nullptr;
  int (*bind_double)(void*, int, double) = nullptr;
  int (*bind_int64)(void*, int, long long) = nullptr;
  int (*bind_text)(void*, int, const char*, int, void (*)(void*)) = nullptr;
  int (*bind_null)(void*, int) = nullptr;
  int (*step)(void*) = nullptr;
  int (*reset)(void*) = nullptr;
  int (*finalize)(void*) = nullptr;
  const char* (*errmsg)(void*) = nullptr;
  bool ok = false;
};

template <typename F>
bool sym(void* h, const char* name, F& fn) {
  fn = reinterpret_cast<F>(dlsym(h, name));
  return fn != nullptr;
}

const Sqlite& sqlite() {
  static Sqlite s = [] {
    Sqlite q;
    for (const char* name : {"libsqlite3.so.0", "libsqlite3.so"}) {
      q.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (q.h) break;
    }
    if (!q.h) return q;
    q.ok = sym(q.h, "sqlite3_open", q.open) && sym(q.h, "sqlite3_close", q.close) && sym(q.h, "sqlite3_exec", q.exec) &&
           sym(q.h, "sqlite3_prepare_v2", q.prepare) && sym(q.h, "sqlite3_bind_blob", q.bind_blob) &&
           sym(q.h, "sqlite3_bind_double", q.bind_double) && sym(q.h, "sqlite3_bind_int64", q.bind_int64) &&
           sym(q.h, "sqlite3_bind_text", q.bind_text) && sym(q.h, "sqlite3_bind_null", q.bind_null) &&
           sym(q.h, "sqlite3_step", q.step) && sym(q.h, "sqlite3_reset", q.reset) &&
           sym(q.h, "sqlite3_finalize", q.finalize) && sym(q.h, "sqlite3_errmsg", q.errmsg);
    return q;
  }();
  return s;
}

constexpr int kSqliteOk = 0, kSqliteDone = 101;

std::string quoted(const char* name) {   // "name" with embedded quotes doubled
  std::string q = "\"";
  for (const char* p = name; *p; ++p) { if (*p == '"') q += '"'; q += *p; }
  return q + "\"";
}

}  // namespace

// Appends n_rings polygon features to table `layer` of the GeoPackage at `path` (tables already created).
//   verts (V,2) f64, ring_off (n_rings+1) i64: ring i = verts[ring_off[i] : ring_off[i+1]] (closed ring as stored);
//   columns: n_cols names / types (0 = float64 array, 1 = int64 array, 2 = text: UTF-8 bytes + (n_rings+1) offsets);
//   col_data[c]: the array (or the bytes), col_text_off[c]: the offsets of a text column (else ignored).
// A NaN of a float column is stored as NULL (what the Python writer does with None).
extern "C" int td_gpkg_append(const char* path, const char* layer, int epsg, const double* verts,
                              const long long* ring_off, long long n_rings, int n_cols, const char* const* col_names,
                              const int* col_types, const void* const* col_data,
                              const long long* const* col_text_off) {
  TD_ARG(path && layer && ring_off && n_rings >= 0 && n_cols >= 0);
  TD_ARG(n_cols == 0 || (col_names && col_types && col_data));
  const Sqlite& q = sqlite();
  if (!q.ok) { td_set_error("td_gpkg_append: libsqlite3.so.0 could not be loaded"); return TD_ERR_UNSUPPORTED; }
  void* db = nullptr;
  if (q.open(path, &db) != kSqliteOk) {
    td_set_error("td_gpkg_append: cannot open %s: %s", path, db ? q.errmsg(db) : "out of memory");
    if (db) q.close(db);
    return TD_ERR_ARG;
  }
  auto fail = [&](const char* what) {
    td_set_error("td_gpkg_append: %s: %s", what, q.errmsg(db));
    q.exec(db, "ROLLBACK", nullptr, nullptr, nullptr);
    q.close(db);
    return TD_ERR_ARG;
  };
  std::string sql = "INSERT INTO " + quoted(layer) + " (geom";
  for (int c = 0; c < n_cols; ++c) sql += ", " + quoted(col_names[c]);
  sql += ") VALUES (?";
  for (int c = 0; c < n_cols; ++c) sql += ",?";
  sql += ")";
  if (q.exec(db, "PRAGMA synchronous = OFF; BEGIN", nullptr, nullptr, nullptr) != kSqliteOk) return fail("BEGIN");
  void* st = nullptr;
  if (q.prepare(db, sql.c_str(), -1, &st, nullptr) != kSqliteOk) return fail("prepare");
  std::vector<unsigned char> blob;
  // GeoPackageBinary header: "GP", version 0, flags 0x03 (little endian, envelope [minx, maxx, miny, maxy])
  unsigned char head[8] = {'G', 'P', 0, 0x03, 0, 0, 0, 0};
  const int32_t srs = epsg;
  std::memcpy(head + 4, &srs, 4);
  // an empty ring: flags 0x11 (empty geometry, no envelope) + WKB polygon with 0 rings
  int rc = TD_OK;
  for (long long i = 0; i < n_rings && rc == TD_OK; ++i) {
    const long long o = ring_off[i], k = ring_off[i + 1] - o;
    if (k < 0 || (k > 0 && !verts)) { td_set_error("td_gpkg_append: ring offsets are not ascending"); rc = TD_ERR_ARG; break; }
    if (k == 0) {
      blob.resize(8 + 9);
      std::memcpy(blob.data(), head, 8);
      blob[3] = 0x11;
      const unsigned char wkb[9] = {1, 3, 0, 0, 0, 0, 0, 0, 0};
      std::memcpy(blob.data() + 8, wkb, 9);
    } else {
      blob.resize(8 + 32 + 13 + 16 * (size_t)k);
      std::memcpy(blob.data(), head, 8);
      const double* p = verts + 2 * o;
      double env[4] = {p[0], p[0], p[1], p[1]};
      for (long long v = 1; v < k; ++v) {
        const double x = p[2 * v], y = p[2 * v + 1];
        // NumPy's minimum / maximum propagate NaN; crown coordinates are finite, plain comparisons suffice
        if (x < env[0]) env[0] = x;
        if (x > env[1]) env[1] = x;
        if (y < env[2]) env[2] = y;
        if (y > env[3]) env[3] = y;
      }
      std::memcpy(blob.data() + 8, env, 32);
      unsigned char* w = blob.data() + 40;
      w[0] = 1;
      const uint32_t gtype = 3, nr = 1, np = (uint32_t)k;
      std::memcpy(w + 1, &gtype, 4); std::memcpy(w + 5, &nr, 4); std::memcpy(w + 9, &np, 4);
      std::memcpy(w + 13, p, 16 * (size_t)k);
    }
    bool ok = q.bind_blob(st, 1, blob.data(), (int)blob.size(), nullptr) == kSqliteOk;   // SQLITE_STATIC: stepped below
    for (int c = 0; c < n_cols && ok; ++c) {
      switch (col_types[c]) {
        case 0: {
          const double v = static_cast<const double*>(col_data[c])[i];
          ok = (v != v ? q.bind_null(st, c + 2) : q.bind_double(st, c + 2, v)) == kSqliteOk;
          break;
        }
        case 1: ok = q.bind_int64(st, c + 2, static_cast<const long long*>(col_data[c])[i]) == kSqliteOk; break;
        case 2: {
          const long long* to = col_text_off ? col_text_off[c] : nullptr;
          if (!to) { ok = false; break; }
          ok = q.bind_text(st, c + 2, static_cast<const char*>(col_data[c]) + to[i], (int)(to[i + 1] - to[i]), nullptr) ==
               kSqliteOk;
          break;
        }
        default: ok = false;
      }
    }
    if (!ok || q.step(st) != kSqliteDone) { td_set_error("td_gpkg_append: row %lld: %s", i, q.errmsg(db)); rc = TD_ERR_ARG; break; }
    q.reset(st);
  }
  q.finalize(st);
  if (rc != TD_OK) {
    q.exec(db, "ROLLBACK", nullptr, nullptr, nullptr);
    q.close(db);
    return rc;
  }
  if (q.exec(db, "COMMIT", nullptr, nullptr, nullptr) != kSqliteOk) return fail("COMMIT");
  q.close(db);
  return TD_OK;
}
