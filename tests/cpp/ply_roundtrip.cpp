// reads a PLY with slam-envire_b200/io/ply.hpp and writes it back (CPU only): tests/test_io_and_drivers.py checks the bytes
#include <cstdio>
#include "io/ply.hpp"
int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    envire::Pointcloud pc;
    const size_t n = envire::io::readPly(argv[1], pc);
    envire::io::writePly(argv[2], pc, argc > 3 ? atoi(argv[3]) != 0 : true);
    printf("%zu %d\n", n, pc.hasData(envire::Pointcloud::VERTEX_NORMAL) ? 1 : 0);
    return 0;
}
