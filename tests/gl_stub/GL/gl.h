/* TEST INFRASTRUCTURE: the handful of OpenGL typedefs cuda_gl_interop.h needs, so that
 * simplehydrology_b200/host/shx_gl.hpp can be compile-checked on a machine without GL headers
 * (tests/test_abi.py).  A real build uses the system's <GL/gl.h>. */
#ifndef SHX_TEST_GL_STUB_H
#define SHX_TEST_GL_STUB_H
typedef unsigned int GLuint;
typedef unsigned int GLenum;
#endif
