// obj_writer.cpp -- .obj export with the reference's exact file bytes.
//
// Reference: File_output::file_write_obj, src/File_output.cu:5-81.  Same observable behaviour:
//   * every vertex is quantised  v = float(int(p * 1000) * 0.001)   (:26-28)
//   * vertices are welded on the quantised triple; first occurrence defines the 1-based id (:30-48)
//   * a face is dropped when two of its ids coincide or when the same ordered id triple was
//     already written (:63-75); winding is flipped on output  " f  a c b" (:72)
//   * numbers are printed with iostream defaults (== "%g")
// The reference welds with std::map<std::vector<float>,int> (a heap allocation and an O(log V)
// lexicographic compare per vertex); this writer uses an open-addressing hash on the quantised
// bit patterns and a single output buffer, which is what makes 10^7..10^8 vertices practical.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace gcb {

namespace {
struct Key { uint32_t a, b, c; };
inline uint64_t mix(uint32_t a, uint32_t b, uint32_t c) {
    uint64_t h = a * 0x9E3779B97F4A7C15ull;
    h ^= (h >> 29);
    h += b * 0xBF58476D1CE4E5B9ull;
    h ^= (h >> 32);
    h += c * 0x94D049BB133111EBull;
    h ^= (h >> 31);
    return h * 0xD6E8FEB86659FD93ull;
}
struct Table {  // key -> 1-based id, 0 = empty
    std::vector<Key> keys;
    std::vector<uint32_t> ids;
    uint64_t mask;
    explicit Table(size_t n) {
        size_t cap = 16;
        while (cap < n * 2) cap <<= 1;
        keys.resize(cap);
        ids.assign(cap, 0);
        mask = cap - 1;
    }
    // returns existing id, or inserts new_id and returns 0
    uint32_t find_or_insert(Key k, uint32_t new_id) {
        uint64_t i = (mix(k.a, k.b, k.c) >> 7) & mask;
        for (;;) {
            if (ids[i] == 0) { keys[i] = k; ids[i] = new_id; return 0; }
            if (keys[i].a == k.a && keys[i].b == k.b && keys[i].c == k.c) return ids[i];
            i = (i + 1) & mask;
        }
    }
};
inline uint32_t fbits(float f) {
    if (f == 0.0f) f = 0.0f;  // the reference's map treats -0 and +0 as the same key
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
inline void put_g(std::string& s, float v) {
    char buf[40];
    int n = snprintf(buf, sizeof buf, "%g", (double)v);
    s.append(buf, (size_t)n);
}
} // namespace

int write_obj_host(const float* pos4, unsigned int total_verts, const char* filename) {
    FILE* f = fopen(filename, "wb");
    if (!f) return 1;
    std::string out;
    out.reserve((size_t)total_verts * 24 + 64);
    out += "##Sample latttice new Obj \n";
    out += "o Solid \n";
    Table weld(total_verts);
    std::vector<uint32_t> faces((size_t)total_verts);
    uint32_t index = 0;
    for (unsigned int i = 0; i < total_verts; ++i) {
        const float* p = pos4 + 4 * (size_t)i;
        const float vx = int(p[0] * 1000) * 0.001, vy = int(p[1] * 1000) * 0.001, vz = int(p[2] * 1000) * 0.001;
        const Key k{fbits(vx), fbits(vy), fbits(vz)};
        uint32_t id = weld.find_or_insert(k, index + 1);
        if (id == 0) {
            id = ++index;
            out += "v ";
            put_g(out, vx); out += ' ';
            put_g(out, vy); out += ' ';
            put_g(out, vz); out += '\n';
        }
        faces[i] = id;
    }
    out += "\n\n";
    Table fseen(total_verts / 3 + 1);
    char buf[64];
    for (size_t i = 0; i + 2 < faces.size(); i += 3) {
        const uint32_t a = faces[i], b = faces[i + 1], c = faces[i + 2];
        if (a != b && a != c && b != c) {
            if (fseen.find_or_insert(Key{a, b, c}, 1) == 0) {
                int n = snprintf(buf, sizeof buf, " f  %u %u %u\n", a, c, b);
                out.append(buf, (size_t)n);
            }
        }
    }
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    return (fclose(f) == 0 && ok) ? 0 : 1;
}

} // namespace gcb
