// TEST INFRASTRUCTURE.  Stand-in for <loom/protobuf_stream.hpp> (which needs libprotobuf's zero-copy streams): InFile
// and OutFile over named in-memory lists of rows, so that Differ::add_rows / compress_rows -- which take file names --
// run unmodified.  The harness fills `memory_files()["name"]` and reads the output list back.
#pragma once
#include <map>
#include <string>
#include <vector>
#include <loom/common.hpp>
#include <loom/schema.pb.h>
namespace loom { namespace protobuf {
inline std::map<std::string, std::vector< ::protobuf::loom::Row>> &memory_files() {
    static std::map<std::string, std::vector< ::protobuf::loom::Row>> files;
    return files;
}
class InFile {
    std::string name_;
    size_t position_ = 0;

public:
    explicit InFile(const char *filename) : name_(filename) {}
    bool is_file() const { return true; }
    const char *filename() const { return name_.c_str(); }
    template <class Message> void read(Message &) {}                                // single-message files: never read by the harnesses
    template <class Message> bool try_read_stream(Message &) { return false; }      // only rows live in memory files
    bool try_read_stream(::protobuf::loom::Row &row) {
        auto &rows = memory_files()[name_];
        if (position_ >= rows.size()) return false;
        row = rows[position_++];
        return true;
    }
};
class OutFile {
    std::string name_;

public:
    explicit OutFile(const char *filename) : name_(filename) { memory_files()[name_].clear(); }
    template <class Message> void write(const Message &) {}
    template <class Message> void write_stream(const Message &) {}
    void write_stream(const ::protobuf::loom::Row &row) { memory_files()[name_].push_back(row); }
};
}  // namespace protobuf
template <class Message> inline std::vector<Message> protobuf_stream_load(const char *) { return std::vector<Message>(); }
}  // namespace loom
