/* TEST INFRASTRUCTURE ONLY -- headless stand-in for <GL/glew.h>.
 *
 * Declares exactly the OpenGL names the reference's main.cpp / render.cpp use,
 * so both compile unmodified from /root/reference.  The definitions live in
 * oracle/ref_oracle.cpp: they are a *recording fake* -- glBufferData keeps the
 * bytes the reference uploads (its patch vertex grid and triangle-strip index
 * buffer, main.cpp:479-480), glTexImage2D keeps every generated height map
 * (main.cpp:245 -> render.cpp:426), glUniform*fv keeps the last per-quad
 * uniforms and glDrawElements counts draws.  No rendering happens.
 */
#ifndef PLANET_ORACLE_GLEW_STUB_H
#define PLANET_ORACLE_GLEW_STUB_H

#include <cstdint>
#include <cstddef>

typedef unsigned int  GLenum;
typedef unsigned int  GLuint;
typedef int           GLint;
typedef int           GLsizei;
typedef unsigned char GLboolean;
typedef unsigned int  GLbitfield;
typedef float         GLfloat;
typedef char          GLchar;
typedef ptrdiff_t     GLsizeiptr;
typedef void          GLvoid;

#define GL_FALSE 0
#define GL_TRUE  1
#define GL_NO_ERROR 0
#define GL_INVALID_ENUM 0x0500
#define GL_INVALID_VALUE 0x0501
#define GL_INVALID_OPERATION 0x0502
#define GL_STACK_OVERFLOW 0x0503
#define GL_STACK_UNDERFLOW 0x0504
#define GL_OUT_OF_MEMORY 0x0505
#define GL_INVALID_FRAMEBUFFER_OPERATION 0x0506
#define GL_DEPTH_BUFFER_BIT 0x00000100
#define GL_COLOR_BUFFER_BIT 0x00004000
#define GL_TRIANGLE_STRIP 0x0005
#define GL_LEQUAL 0x0203
#define GL_FRONT_AND_BACK 0x0408
#define GL_BACK 0x0405
#define GL_CW 0x0900
#define GL_CULL_FACE 0x0B44
#define GL_DEPTH_TEST 0x0B71
#define GL_TEXTURE_2D 0x0DE1
#define GL_UNSIGNED_INT 0x1405
#define GL_FLOAT 0x1406
#define GL_RED 0x1903
#define GL_LINE 0x1B01
#define GL_FILL 0x1B02
#define GL_NEAREST 0x2600
#define GL_LINEAR 0x2601
#define GL_NEAREST_MIPMAP_NEAREST 0x2700
#define GL_LINEAR_MIPMAP_NEAREST 0x2701
#define GL_LINEAR_MIPMAP_LINEAR 0x2703
#define GL_TEXTURE_MAG_FILTER 0x2800
#define GL_TEXTURE_MIN_FILTER 0x2801
#define GL_TEXTURE_WRAP_S 0x2802
#define GL_TEXTURE_WRAP_T 0x2803
#define GL_REPEAT 0x2901
#define GL_CLAMP_TO_EDGE 0x812F
#define GL_R32F 0x822E
#define GL_MIRRORED_REPEAT 0x8370
#define GL_TEXTURE0 0x84C0
#define GL_ARRAY_BUFFER 0x8892
#define GL_ELEMENT_ARRAY_BUFFER 0x8893
#define GL_STATIC_DRAW 0x88E4
#define GL_FRAGMENT_SHADER 0x8B30
#define GL_VERTEX_SHADER 0x8B31
#define GL_COMPILE_STATUS 0x8B81
#define GL_LINK_STATUS 0x8B82

#define GLEW_OK 0
#define GLEW_VERSION_3_1 1
extern GLboolean glewExperimental;
GLenum glewInit();
const unsigned char *glewGetErrorString(GLenum);

void glActiveTexture(GLenum);
void glAttachShader(GLuint, GLuint);
void glBindAttribLocation(GLuint, GLuint, const GLchar *);
void glBindBuffer(GLenum, GLuint);
void glBindTexture(GLenum, GLuint);
void glBindVertexArray(GLuint);
void glBufferData(GLenum, GLsizeiptr, const void *, GLenum);
void glClear(GLbitfield);
void glCompileShader(GLuint);
GLuint glCreateProgram();
GLuint glCreateShader(GLenum);
void glCullFace(GLenum);
void glDeleteProgram(GLuint);
void glDeleteShader(GLuint);
void glDeleteTextures(GLsizei, const GLuint *);
void glDepthFunc(GLenum);
void glDetachShader(GLuint, GLuint);
void glDrawArrays(GLenum, GLint, GLsizei);
void glDrawElements(GLenum, GLsizei, GLenum, const void *);
void glEnable(GLenum);
void glEnableVertexAttribArray(GLuint);
void glFrontFace(GLenum);
void glGenBuffers(GLsizei, GLuint *);
void glGenTextures(GLsizei, GLuint *);
void glGenVertexArrays(GLsizei, GLuint *);
GLenum glGetError();
void glGetProgramInfoLog(GLuint, GLsizei, GLsizei *, GLchar *);
void glGetProgramiv(GLuint, GLenum, GLint *);
void glGetShaderInfoLog(GLuint, GLsizei, GLsizei *, GLchar *);
void glGetShaderiv(GLuint, GLenum, GLint *);
GLint glGetUniformLocation(GLuint, const GLchar *);
void glLinkProgram(GLuint);
void glPolygonMode(GLenum, GLenum);
void glShaderSource(GLuint, GLsizei, const GLchar *const *, const GLint *);
void glTexImage2D(GLenum, GLint, GLint, GLsizei, GLsizei, GLint, GLenum, GLenum, const void *);
void glTexParameteri(GLenum, GLenum, GLint);
void glUniform1fv(GLint, GLsizei, const GLfloat *);
void glUniform1i(GLint, GLint);
void glUniform2fv(GLint, GLsizei, const GLfloat *);
void glUniform3fv(GLint, GLsizei, const GLfloat *);
void glUniformMatrix4fv(GLint, GLsizei, GLboolean, const GLfloat *);
void glUseProgram(GLuint);
void glVertexAttribPointer(GLuint, GLint, GLenum, GLboolean, GLsizei, const void *);
void glViewport(GLint, GLint, GLsizei, GLsizei);

#endif
