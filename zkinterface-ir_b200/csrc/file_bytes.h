// The bytes of one `.sieve` file (consumers/source.rs:121-159 reads a file message by message into fresh Vecs):
// regular files are mapped (no copy: the FlatBuffers reader walks the page cache), stdin and pipes are read.
#pragma once
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

namespace zkb {

class FileBytes {
public:
    const uint8_t* data = nullptr;
    size_t size = 0;

    FileBytes() {}
    FileBytes(const FileBytes&) = delete;
    FileBytes& operator=(const FileBytes&) = delete;
    ~FileBytes() {
        if (mapped_) munmap(mapped_, size);
    }
    // "-" is stdin
    bool open(const std::string& path) {
        if (path != "-") {
            int fd = ::open(path.c_str(), O_RDONLY);
            if (fd < 0) return false;
            struct stat st;
            if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) {
                if (st.st_size == 0) {
                    close(fd);
                    return true;
                }
                void* p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                if (p != MAP_FAILED) {
                    madvise(p, (size_t)st.st_size, MADV_SEQUENTIAL | MADV_WILLNEED);
                    mapped_ = p;
                    data = (const uint8_t*)p;
                    size = (size_t)st.st_size;
                    close(fd);
                    return true;
                }
            }
            bool ok = slurp(fd);
            close(fd);
            return ok;
        }
        return slurp(0);
    }

private:
    bool slurp(int fd) {
        size_t used = 0;
        buf_.resize(1 << 20);
        for (;;) {
            if (used == buf_.size()) buf_.resize(buf_.size() * 2);
            ssize_t got = read(fd, buf_.data() + used, buf_.size() - used);
            if (got < 0) return false;
            if (got == 0) break;
            used += (size_t)got;
        }
        buf_.resize(used);
        data = buf_.data();
        size = used;
        return true;
    }
    void* mapped_ = nullptr;
    std::vector<uint8_t> buf_;
};

}  // namespace zkb
