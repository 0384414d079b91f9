// rmgpu_server -- persistent per-box worker behind the same C ABI (SURVEY 7 "hard part 5", 8f.4).
//
//   rmgpu_server --socket /tmp/rmgpu.sock [--gpus 0,1,2,...]
//
// RepeatMasker builds its engine object in the parent and forks one short-lived child per batch
// (RepeatMasker:602, :1043); a CUDA context cannot cross fork and creating one per child (plus re-reading and
// re-packing the library on every search(), as the reference does with cross_match) would dwarf the search itself.
// With RMGPU_SERVER=<socket> in the environment, rmgpu_search() in the children becomes a thin client: it sends the
// three paths and the parameters, this process runs the search on one of its GPU workers -- each with its own
// rmgpu_session, i.e. matrix and packed library resident in HBM as long as path, size and mtime do not change -- and
// returns the flat hit records.  Batches are independent, so workers never talk to each other (no collective).
#include <signal.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rmgpu.h"
#include "remote.hpp"

using namespace rmgpu;

namespace {

struct Worker {
  int device = 0;
  std::mutex mu;
  rmgpu_session* session = nullptr;
  std::string matrixKey, libraryKey;
  long served = 0;
};

std::vector<Worker*> g_workers;
std::atomic<unsigned> g_next{0};
std::atomic<bool> g_quit{false};
std::string g_socket;
int g_listen = -1;

std::string file_key(const std::string& path) {
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return path + "|missing";
  return path + "|" + std::to_string((long long)st.st_size) + "|" + std::to_string((long long)st.st_mtim.tv_sec) + "." +
         std::to_string((long long)st.st_mtim.tv_nsec);
}

bool read_str(int fd, std::string& s) {
  uint32_t n = 0;
  if (!read_full(fd, &n, 4)) return false;
  if (n == 0xFFFFFFFFu) {
    s.clear();
    return true;
  }
  if (n > (1u << 20)) return false;
  s.resize(n);
  return n == 0 || read_full(fd, &s[0], n);
}

void reply_error(int fd, int status, const std::string& msg) {
  int32_t st = status ? status : 1;
  uint32_t n = (uint32_t)msg.size();
  write_full(fd, &st, 4);
  write_full(fd, &n, 4);
  write_full(fd, msg.data(), n);
}

void handle(int fd) {
  uint32_t magic = 0;
  if (!read_full(fd, &magic, 4)) {
    ::close(fd);
    return;
  }
  if (magic == kWirePing || magic == kWireQuit) {
    int32_t st = 0;
    write_full(fd, &st, 4);
    ::close(fd);
    if (magic == kWireQuit) {
      g_quit = true;
      ::shutdown(g_listen, SHUT_RDWR);
    }
    return;
  }
  rmgpu_params p;
  std::string query, subject, matrix;
  if (magic != kWireSearch || !read_full(fd, &p, sizeof(p)) || !read_str(fd, query) || !read_str(fd, subject) ||
      !read_str(fd, matrix)) {
    reply_error(fd, 1, "rmgpu_server: malformed request");
    ::close(fd);
    return;
  }
  // pick a worker: first idle one starting from a rotating index, else wait for the rotating one
  const unsigned start = g_next++ % g_workers.size();
  Worker* w = nullptr;
  for (size_t i = 0; i < g_workers.size() && !w; i++) {
    Worker* c = g_workers[(start + i) % g_workers.size()];
    if (c->mu.try_lock()) w = c;
  }
  if (!w) {
    w = g_workers[start];
    w->mu.lock();
  }
  std::vector<uint8_t> body;
  int rc = 0;
  std::string err;
  {
    if (!w->session) {
      w->session = rmgpu_session_open(w->device);
      if (!w->session) {
        rc = 1;
        err = rmgpu_last_error();
      }
    }
    if (!rc) {
      const std::string mk = file_key(matrix), lk = file_key(subject);
      if (mk != w->matrixKey) {
        rc = rmgpu_session_set_matrix_file(w->session, matrix.c_str());
        w->matrixKey = rc ? "" : mk;
        w->libraryKey.clear();  // residue codes depend on the matrix alphabet
      }
      if (!rc && lk != w->libraryKey) {
        rc = rmgpu_session_set_library_fasta(w->session, subject.c_str());
        w->libraryKey = rc ? "" : lk;
      }
      if (!rc) rc = rmgpu_session_set_queries_fasta(w->session, query.c_str());
      rmgpu_result* res = nullptr;
      p.device = w->device;
      if (!rc) rc = rmgpu_session_search(w->session, &p, &res);
      if (rc)
        err = rmgpu_last_error();
      else {
        serialize_result(res, body);
        rmgpu_free(res);
        w->served++;
      }
    }
  }
  w->mu.unlock();
  if (rc) {
    reply_error(fd, rc, err);
  } else {
    int32_t st = 0;
    uint32_t zero = 0;
    uint64_t n = body.size();
    write_full(fd, &st, 4);
    write_full(fd, &zero, 4);
    write_full(fd, &n, 8);
    write_full(fd, body.data(), body.size());
  }
  ::close(fd);
}

void on_signal(int) {
  g_quit = true;
  if (g_listen >= 0) ::shutdown(g_listen, SHUT_RDWR);
}

}  // namespace

int main(int argc, char** argv) {
  std::string gpus = "0";
  for (int a = 1; a < argc; a++) {
    if (!strcmp(argv[a], "--socket") && a + 1 < argc)
      g_socket = argv[++a];
    else if (!strcmp(argv[a], "--gpus") && a + 1 < argc)
      gpus = argv[++a];
    else {
      fprintf(stderr, "usage: rmgpu_server --socket PATH [--gpus 0,1,...]\n%s\n", rmgpu_version());
      return 2;
    }
  }
  if (g_socket.empty()) {
    fprintf(stderr, "usage: rmgpu_server --socket PATH [--gpus 0,1,...]\n%s\n", rmgpu_version());
    return 2;
  }
  for (size_t pos = 0; pos < gpus.size();) {
    size_t comma = gpus.find(',', pos);
    if (comma == std::string::npos) comma = gpus.size();
    auto* w = new Worker();
    w->device = atoi(gpus.substr(pos, comma - pos).c_str());
    g_workers.push_back(w);
    pos = comma + 1;
  }
  signal(SIGPIPE, SIG_IGN);
  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);
  ::unlink(g_socket.c_str());
  g_listen = ::socket(AF_UNIX, SOCK_STREAM, 0);
  sockaddr_un sa;
  memset(&sa, 0, sizeof(sa));
  sa.sun_family = AF_UNIX;
  if (g_socket.size() >= sizeof(sa.sun_path)) {
    fprintf(stderr, "rmgpu_server: socket path too long\n");
    return 2;
  }
  strcpy(sa.sun_path, g_socket.c_str());
  if (g_listen < 0 || ::bind(g_listen, (sockaddr*)&sa, sizeof(sa)) != 0 || ::listen(g_listen, 256) != 0) {
    perror("rmgpu_server: bind/listen");
    return 1;
  }
  fprintf(stderr, "rmgpu_server: %s, %zu GPU worker(s), listening on %s\n", rmgpu_version(), g_workers.size(), g_socket.c_str());
  std::vector<std::thread> threads;
  while (!g_quit) {
    int fd = ::accept(g_listen, nullptr, nullptr);
    if (fd < 0) {
      if (g_quit) break;
      continue;
    }
    threads.emplace_back(handle, fd);
    if (threads.size() > 512) {  // reap finished connection threads now and then
      for (auto& t : threads) t.join();
      threads.clear();
    }
  }
  for (auto& t : threads) t.join();
  for (auto* w : g_workers) {
    if (w->session) rmgpu_session_close(w->session);
    fprintf(stderr, "rmgpu_server: GPU %d served %ld searches\n", w->device, w->served);
  }
  ::close(g_listen);
  ::unlink(g_socket.c_str());
  return 0;
}
