// Chunked read of a byte range of a file through two alternating staging slots - the host half of
// b200_table_upload_file (point slices of gnark's ProvingKey dump, backend/groth16/bn254/marshal.go:375-539, go from
// the file to HBM without passing through a Go slice).  Plain POSIX + C++, no CUDA: the copy engine side is the
// caller's `sink`; tests/test_abi.py exercises this file on the CPU through the emulation library.
#pragma once
#include <unistd.h>

#include <cerrno>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>

namespace gb200 {

// Reads [off, off + bytes) of fd, slot_bytes at a time, into slots[0], slots[1], slots[0], ...  Before a slot is
// overwritten `wait(slot)` must return 0 (its previous contents have been consumed); after a chunk has been read
// `sink(slot, data, pos, len)` consumes it (pos = offset inside the range).  Returns 0, or -1 with *err set.
template <class SINK, class WAIT>
int stage_file_region(int fd, uint64_t off, size_t bytes, void* const slots[2], size_t slot_bytes, SINK sink, WAIT wait,
                      std::string* err) {
  if (slot_bytes == 0) { *err = "stage_file_region: empty staging slot"; return -1; }
  size_t pos = 0;
  int slot = 0;
  while (pos < bytes) {
    const size_t len = bytes - pos < slot_bytes ? bytes - pos : slot_bytes;
    if (wait(slot) != 0) { *err = "stage_file_region: staging slot not released"; return -1; }
    size_t got = 0;
    while (got < len) {
      const ssize_t r = pread(fd, (char*)slots[slot] + got, len - got, (off_t)(off + pos + got));
      if (r < 0) {
        if (errno == EINTR) continue;
        *err = std::string("stage_file_region: pread: ") + strerror(errno);
        return -1;
      }
      if (r == 0) { *err = "stage_file_region: file shorter than the requested range"; return -1; }
      got += (size_t)r;
    }
    if (sink(slot, slots[slot], pos, len) != 0) { *err = "stage_file_region: sink failed"; return -1; }
    pos += len;
    slot ^= 1;
  }
  return 0;
}

}  // namespace gb200
