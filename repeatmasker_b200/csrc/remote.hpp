// remote.hpp -- wire format between librmgpu's client side (rmgpu_search with RMGPU_SERVER set) and rmgpu_server.
// Little-endian, length-prefixed, over an AF_UNIX stream socket.  One request per connection.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rmgpu.h"

namespace rmgpu {

constexpr uint32_t kWireSearch = 0x31474D52u;  // "RMG1": params, query path, subject path, matrix path
constexpr uint32_t kWireQuit = 0x51474D52u;    // "RMGQ": ask the server to exit
constexpr uint32_t kWirePing = 0x50474D52u;    // "RMGP": liveness probe, answered with status 0

// client: returns 0 and *out on success; non-zero and err otherwise
int remote_search(const char* socket_path, const rmgpu_params* p, const char* query, const char* subject, const char* matrix,
                  rmgpu_result** out, std::string& err);
int remote_command(const char* socket_path, uint32_t magic, std::string& err);

// (de)serialisation of a result
void serialize_result(const rmgpu_result* r, std::vector<uint8_t>& buf);
rmgpu_result* deserialize_result(const uint8_t* p, size_t n);

bool read_full(int fd, void* buf, size_t n);
bool write_full(int fd, const void* buf, size_t n);

}  // namespace rmgpu
