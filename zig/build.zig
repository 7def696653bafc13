// Patched build.zig for the reference (replaces /build.zig, 19 lines): same executable and run
// step, plus the link of the prebuilt CUDA backend.  NOT TESTED — no zig toolchain in the image.
//
//   1. python -m rayz_b200.build              (nvcc -gencode arch=compute_100a,code=sm_100a -> librayz_cuda.so)
//   2. copy zig/cuda_backend.zig to src/, apply the renderer.zig edit shown in INTEGRATION.md
//   3. zig build -Doptimize=ReleaseFast -Drayz-cuda=/path/to/rayz_b200/lib run -- 1200 out.ppm
const std = @import("std");

pub fn build(b: *std.Build) void {
    const optimize = b.standardOptimizeOption(.{}); // the reference sets none => Debug
    const cuda_lib_dir = b.option([]const u8, "rayz-cuda", "directory holding librayz_cuda.so") orelse "rayz_b200/lib";

    const exe = b.addExecutable(.{
        .name = "rayz",
        .root_source_file = b.path("src/rayz.zig"),
        .target = b.graph.host,
        .optimize = optimize,
    });
    exe.linkLibC();
    exe.addLibraryPath(.{ .cwd_relative = cuda_lib_dir });
    exe.addRPath(.{ .cwd_relative = cuda_lib_dir });
    exe.linkSystemLibrary("rayz_cuda"); // cudart is linked statically inside it; libcuda comes from the driver

    b.installArtifact(exe);

    const run_exe = b.addRunArtifact(exe);
    if (b.args) |args| {
        run_exe.addArgs(args);
    }
    const run_step = b.step("run", "Run the application");
    run_step.dependOn(&run_exe.step);
}
