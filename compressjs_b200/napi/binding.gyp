{
  # node-gyp recipe of the N-API addon.  It links the in-tree libb2bz.so (built by `make -C ../csrc` or
  # __graft_entry__.build()); the addon itself is plain C++ over node_api.h, no CUDA at this level.
  # UNBUILT in this repository's image (no node, no node-gyp, no headers): `npm install` on a box with node >= 12
  # and the CUDA runtime is the first build.
  "targets": [
    {
      "target_name": "b2bz",
      "sources": ["addon.cc"],
      "include_dirs": ["../../include"],
      "defines": ["NAPI_VERSION=6"],
      "cflags_cc": ["-std=c++17", "-O2"],
      "libraries": ["-L<(module_root_dir)/..", "-lb2bz", "-Wl,-rpath,<(module_root_dir)/.."]
    }
  ]
}
