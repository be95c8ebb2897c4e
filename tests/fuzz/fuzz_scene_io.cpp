// fuzz_scene_io.cpp — TEST INFRASTRUCTURE: mutation fuzzing of the native scene loaders and texture decoders
// (chameleonrt_b200/csrc/scene_io.cpp, image_decode.h, jpeg_decode.h, json_reader.h) under AddressSanitizer + UndefinedBehaviorSanitizer. Every seed file is
// mutated (random bytes, bit flips, truncation, "interesting" bytes) and loaded; a load may succeed or throw, but must not
// touch memory it does not own or run into undefined behaviour. Built and driven by tests/test_scene_io_fuzz.py:
//     fuzz_scene_io <seed dir> <work dir> <iterations per seed>
// The product source is included so that its file-local functions can be called directly.
#include "../../chameleonrt_b200/csrc/scene_io.cpp"
#include <cstdio>
#include <random>
static std::vector<uint8_t> read_file(const std::string &p)
{
    std::ifstream in(p.c_str(), std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}
int main(int argc, char **argv)
{
    if (argc < 4) {
        return 2;
    }
    const std::string seed_dir = argv[1], work_dir = argv[2];
    const int iters = atoi(argv[3]);
    const char *names[] = {"a.jpg", "b.jpg", "c.jpg", "a.png", "b.png", "c.png", "a.tga", "b.tga", "a.bmp", "b.bmp", "poly.obj", "h.glb", "h.gltf", "s.crts"};
    std::mt19937 rng(12345);
    int ok = 0, failed = 0;
    for (const char *name : names) {
        const std::vector<uint8_t> seed = read_file(seed_dir + "/" + name);
        const std::string ext = std::string(name).substr(std::string(name).rfind('.'));
        for (int it = 0; it < iters; ++it) {
            std::vector<uint8_t> m = seed;
            const int kind = rng() % 4;
            const int nmut = 1 + rng() % 8;
            for (int k = 0; k < nmut && !m.empty(); ++k) {
                const size_t pos = rng() % m.size();
                if (kind == 0) m[pos] = (uint8_t)rng();
                else if (kind == 1) m[pos] ^= (uint8_t)(1u << (rng() % 8));
                else if (kind == 2) { m.resize(pos); break; }
                else { const uint8_t special[] = {0, 0xff, 0x7f, 0x80, 1, '0', '-', '/', ' ', '\n', '{', '[', '"'}; m[pos] = special[rng() % sizeof(special)]; }
            }
            try {
                if (ext == ".jpg" || ext == ".png" || ext == ".tga" || ext == ".bmp") {
                    std::vector<uint8_t> out;
                    int w, h;
                    decode_image_rgba(m.data(), m.size(), name, out, w, h, it & 1);
                } else {
                    const std::string path = work_dir + "/m" + ext;
                    FILE *f = fopen(path.c_str(), "wb");
                    fwrite(m.data(), 1, m.size(), f);
                    fclose(f);
                    crtio_scene s;
                    if (ext == ".obj") load_obj_impl(path, 2, s);
                    else if (ext == ".crts") load_crts_impl(path, 2, s);
                    else load_gltf_impl(path, 2, s);
                }
                ++ok;
            } catch (const std::exception &) {
                ++failed;
            }
        }
        printf("%s done (ok %d, rejected %d)\n", name, ok, failed);
        fflush(stdout);
    }
}
