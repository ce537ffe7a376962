// shx CUDA-GL interop (BASELINE configs[4]: "CUDA-GL interop update of the vertex pool"), compile-time option:
// define SHX_WITH_GL and include the GL headers before this file.
//
// The reference keeps its terrain mesh in ONE OpenGL buffer, created with glBufferStorage and persistently mapped for
// the CPU (source/vertexpool.h:155-172); every frame `updatenode` (source/cellpool.h:286-305, called from
// SimpleHydrology.cpp:322-324) rewrites the 48-byte Vertex records of all cells through that mapping.  With the map on
// the GPU the same records are produced by shx_vertex_fill; registering the pool's buffer object with CUDA lets that
// kernel write straight into the VBO the renderer draws from, so no vertex (and no height) crosses PCIe:
//
//   shx::GLBuffer vbo(vertexpool.vbo);                  // once, after Vertexpool::reserve (vertexpool.h:155)
//   shx::GLBuffer trees(modelbuf.index);                // the instance buffer of the tree particle system (optional)
//   ...per frame, instead of the updatenode loop and the treemodels loop (SimpleHydrology.cpp:322-335):
//   { auto m = vbo.map();   bridge.update_vertices_device(m.as<float>()); }        // unmapped when m goes out of scope
//   { auto m = trees.map(); treeparticle.SIZE = bridge.tree_models_device(m.as<float>(), m.bytes / 64); }
//
// The pool's sections are reserved node by node in pool order (cellpool.h:327-336), which is the order
// shx_vertex_fill writes.  The buffer must not be mapped by GL for the CPU while CUDA has it mapped: create it with
// glBufferData / glBufferStorage WITHOUT GL_MAP_PERSISTENT_BIT when the host no longer writes vertices.
// Not compiled in this repository's own builds (the image has no GL headers or context); tests/test_abi.py
// compile-checks it against the CUDA toolkit's cuda_gl_interop.h with a two-typedef GL stub.
#pragma once
#ifdef SHX_WITH_GL

#include <cstddef>
#include <stdexcept>
#include <string>

#include <cuda_runtime_api.h>
#include <cuda_gl_interop.h>

namespace shx {

class GLBuffer {
 public:
  // `buffer`: the OpenGL buffer object name; write_discard: CUDA overwrites the whole buffer every frame
  explicit GLBuffer(unsigned buffer, bool write_discard = true) {
    gl_check(cudaGraphicsGLRegisterBuffer(&res_, (GLuint)buffer,
                                          write_discard ? cudaGraphicsRegisterFlagsWriteDiscard : cudaGraphicsRegisterFlagsNone),
             "cudaGraphicsGLRegisterBuffer");
  }
  ~GLBuffer() {
    if (res_) cudaGraphicsUnregisterResource(res_);
  }
  GLBuffer(const GLBuffer&) = delete;
  GLBuffer& operator=(const GLBuffer&) = delete;

  // RAII mapping: the device pointer is valid until the object dies; work queued on `stream` before that is ordered
  // before GL's next use of the buffer
  class Mapping {
   public:
    Mapping(cudaGraphicsResource_t res, cudaStream_t stream) : res_(res), stream_(stream) {
      gl_check(cudaGraphicsMapResources(1, &res_, stream_), "cudaGraphicsMapResources");
      gl_check(cudaGraphicsResourceGetMappedPointer(&ptr, &bytes, res_), "cudaGraphicsResourceGetMappedPointer");
    }
    ~Mapping() { cudaGraphicsUnmapResources(1, &res_, stream_); }
    Mapping(const Mapping&) = delete;
    Mapping& operator=(const Mapping&) = delete;
    template <class T>
    T* as() const { return static_cast<T*>(ptr); }
    void* ptr = nullptr;
    size_t bytes = 0;

   private:
    cudaGraphicsResource_t res_;
    cudaStream_t stream_;
  };
  Mapping map(cudaStream_t stream = nullptr) { return Mapping(res_, stream); }

 private:
  static void gl_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
  }
  cudaGraphicsResource_t res_ = nullptr;
};

}  // namespace shx

#endif  // SHX_WITH_GL
