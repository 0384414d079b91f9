// remote.cpp -- client side of the persistent-worker protocol and the (de)serialisation of results.
#include "remote.hpp"

#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>

#include "host.hpp"

namespace rmgpu {

bool read_full(int fd, void* buf, size_t n) {
  uint8_t* p = (uint8_t*)buf;
  while (n) {
    ssize_t k = ::read(fd, p, n);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) return false;
    p += k;
    n -= (size_t)k;
  }
  return true;
}

bool write_full(int fd, const void* buf, size_t n) {
  const uint8_t* p = (const uint8_t*)buf;
  while (n) {
    ssize_t k = ::write(fd, p, n);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) return false;
    p += k;
    n -= (size_t)k;
  }
  return true;
}

namespace {

void put(std::vector<uint8_t>& b, const void* p, size_t n) { b.insert(b.end(), (const uint8_t*)p, (const uint8_t*)p + n); }
void put_u32(std::vector<uint8_t>& b, uint32_t v) { put(b, &v, 4); }
void put_str(std::vector<uint8_t>& b, const char* s) {
  uint32_t n = s ? (uint32_t)strlen(s) : 0xFFFFFFFFu;
  put_u32(b, n);
  if (s) put(b, s, n);
}

struct Reader {
  const uint8_t* p;
  size_t n;
  bool ok = true;
  void get(void* dst, size_t k) {
    if (k > n) {
      ok = false;
      memset(dst, 0, k);
      return;
    }
    memcpy(dst, p, k);
    p += k;
    n -= k;
  }
  uint32_t u32() {
    uint32_t v = 0;
    get(&v, 4);
    return v;
  }
  const char* str(ResultArena& a) {
    uint32_t k = u32();
    if (k == 0xFFFFFFFFu) return nullptr;
    if (k > n) {
      ok = false;
      return nullptr;
    }
    const char* s = a.dup((const char*)p, k);
    p += k;
    n -= k;
    return s;
  }
};

// the server does not share the client's working directory
std::string absolutize(const char* path) {
  if (!path || path[0] == '/' || path[0] == 0) return path ? path : "";
  char cwd[4096];
  if (!getcwd(cwd, sizeof(cwd))) return path;
  return std::string(cwd) + "/" + path;
}

int connect_unix(const char* path, std::string& err) {
  int fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
  if (fd < 0) {
    err = std::string("socket(): ") + strerror(errno);
    return -1;
  }
  sockaddr_un sa;
  memset(&sa, 0, sizeof(sa));
  sa.sun_family = AF_UNIX;
  if (strlen(path) >= sizeof(sa.sun_path)) {
    err = "RMGPU_SERVER socket path too long";
    ::close(fd);
    return -1;
  }
  strcpy(sa.sun_path, path);
  if (::connect(fd, (sockaddr*)&sa, sizeof(sa)) != 0) {
    err = std::string("cannot reach rmgpu_server at ") + path + ": " + strerror(errno);
    ::close(fd);
    return -1;
  }
  return fd;
}

}  // namespace

void serialize_result(const rmgpu_result* r, std::vector<uint8_t>& b) {
  put(b, &r->stats, sizeof(r->stats));
  put_u32(b, (uint32_t)r->n_hits);
  for (int i = 0; i < r->n_hits; i++) {
    rmgpu_hit h = r->hits[i];
    const char *qn = h.query_name, *sn = h.subject_name, *qa = h.qaln, *sa = h.saln;
    h.query_name = h.subject_name = h.qaln = h.saln = nullptr;  // pointers do not travel
    put(b, &h, sizeof(h));
    put_str(b, qn);
    put_str(b, sn);
    put_str(b, qa);
    put_str(b, sa);
  }
}

rmgpu_result* deserialize_result(const uint8_t* p, size_t n) {
  Reader rd{p, n};
  auto* res = new rmgpu_result();
  memset(res, 0, sizeof(*res));
  auto* arena = new ResultArena();
  res->opaque = arena;
  rd.get(&res->stats, sizeof(res->stats));
  uint32_t nh = rd.u32();
  if (!rd.ok || (size_t)nh * sizeof(rmgpu_hit) > n) {
    delete arena;
    delete res;
    return nullptr;
  }
  res->n_hits = (int32_t)nh;
  res->hits = nh ? new rmgpu_hit[nh] : nullptr;
  for (uint32_t i = 0; i < nh && rd.ok; i++) {
    rd.get(&res->hits[i], sizeof(rmgpu_hit));
    res->hits[i].query_name = rd.str(*arena);
    res->hits[i].subject_name = rd.str(*arena);
    res->hits[i].qaln = rd.str(*arena);
    res->hits[i].saln = rd.str(*arena);
  }
  if (!rd.ok) {
    delete[] res->hits;
    delete arena;
    delete res;
    return nullptr;
  }
  return res;
}

int remote_command(const char* socket_path, uint32_t magic, std::string& err) {
  int fd = connect_unix(socket_path, err);
  if (fd < 0) return 1;
  int32_t status = 1;
  bool ok = write_full(fd, &magic, 4) && read_full(fd, &status, 4);
  ::close(fd);
  if (!ok) {
    err = "rmgpu_server closed the connection";
    return 1;
  }
  return status;
}

int remote_search(const char* socket_path, const rmgpu_params* p, const char* query, const char* subject, const char* matrix,
                  rmgpu_result** out, std::string& err) {
  *out = nullptr;
  int fd = connect_unix(socket_path, err);
  if (fd < 0) return 1;
  std::vector<uint8_t> req;
  put_u32(req, kWireSearch);
  put(req, p, sizeof(*p));
  put_str(req, absolutize(query).c_str());
  put_str(req, absolutize(subject).c_str());
  put_str(req, absolutize(matrix).c_str());
  int32_t status = 1;
  uint32_t elen = 0;
  uint64_t blen = 0;
  bool ok = write_full(fd, req.data(), req.size()) && read_full(fd, &status, 4) && read_full(fd, &elen, 4);
  std::string emsg;
  if (ok && elen) {
    emsg.resize(elen);
    ok = read_full(fd, &emsg[0], elen);
  }
  std::vector<uint8_t> body;
  if (ok && status == 0) {
    ok = read_full(fd, &blen, 8);
    if (ok) {
      body.resize((size_t)blen);
      ok = read_full(fd, body.data(), body.size());
    }
  }
  ::close(fd);
  if (!ok) {
    err = "rmgpu_server closed the connection";
    return 1;
  }
  if (status != 0) {
    err = emsg.empty() ? "rmgpu_server reported an error" : emsg;
    return status;
  }
  *out = deserialize_result(body.data(), body.size());
  if (!*out) {
    err = "malformed reply from rmgpu_server";
    return 1;
  }
  return 0;
}

}  // namespace rmgpu
