"""Dependency-free VTK XML ImageData (.vti) writer for the coarse-grained fields -- the Python twin of
lgca_b200/host/lgca_io_vti.cpp (same layout: raw appended Float32 arrays, header_type UInt64), standing in for the
reference's VTK-based IoVti::write (src/lgca_io_vti.cpp:113-144: `mean_res_<step>.vti` with point data
"Mean density" and "Mean momentum").  Host-side I/O only."""
import struct


def write_mean_vti(path, coarse_dim_x, coarse_dim_y, mean_density, mean_momentum, origin_y=0):
    nx, ny = int(coarse_dim_x), int(coarse_dim_y)
    assert mean_density.size == nx * ny and mean_momentum.size == 2 * nx * ny
    head = ('<?xml version="1.0"?>\n<VTKFile type="ImageData" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n'
            '  <ImageData WholeExtent="0 %d 0 %d 0 0" Origin="0 %d 0" Spacing="1 1 1">\n    <Piece Extent="0 %d 0 %d 0 0">\n'
            '      <PointData Scalars="Mean density">\n'
            '        <DataArray type="Float32" Name="Mean density" NumberOfComponents="1" format="appended" offset="0"/>\n'
            '        <DataArray type="Float32" Name="Mean momentum" NumberOfComponents="2" format="appended" offset="%d"/>\n'
            '      </PointData>\n    </Piece>\n  </ImageData>\n  <AppendedData encoding="raw">\n   _'
            % (nx - 1, ny - 1, origin_y, nx - 1, ny - 1, 8 + mean_density.nbytes))
    with open(path, "wb") as f:
        f.write(head.encode())
        f.write(struct.pack("<Q", mean_density.nbytes))
        f.write(memoryview(mean_density).cast("B"))
        f.write(struct.pack("<Q", mean_momentum.nbytes))
        f.write(memoryview(mean_momentum).cast("B"))
        f.write(b"\n  </AppendedData>\n</VTKFile>\n")
