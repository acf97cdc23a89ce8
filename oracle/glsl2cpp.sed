# glsl2cpp.sed -- syntactic GLSL 4.60 -> C++ mapping for the reference's raytracer.glsl.
# TEST INFRASTRUCTURE ONLY (see oracle/ref_raytracer.cpp, oracle/Makefile).  Applied with
# `sed -E -f`; the output goes to oracle/_ref/ (git-ignored) and is deleted after the compile.
# Nothing here changes an expression: only qualifiers GLSL has and C++ has not.

# preprocessor / interface qualifiers
/^#version/d
/^layout *\(local_size/d
# storage blocks:  layout (binding = N, std430) readonly buffer X { T name[]; };  ->  T* name;
/^layout.*buffer/,/^\};/{
  /^layout/d
  /^\};/d
  s/^ *([A-Za-z_0-9]+) +([A-Za-z_0-9]+)\[\];/\1* \2;/
}
# image binding:  layout(rgba32f, binding = 0) uniform image2D oImage;  ->  image2D oImage;
s/^layout *\([^)]*\) *uniform /uniform /
# uniforms become namespace-scope variables the driver sets
s/^uniform //
# reference parameters
s/\binout +([A-Za-z_0-9]+) +/\1\& /g
# swizzles are member functions in GLM (both swizzle modes accept the call form)
s/\.xyz\b/.xyz()/g
s/\.xy\b/.xy()/g
# an uninitialised `Hit x;` gets the value the build pins for "undefined" (ref_raytracer.cpp header, Q6)
s/^( *)Hit ([A-Za-z_0-9]+);/\1Hit \2 = RTR_UNDEFINED_HIT;/
# entry point
s/^void main\(\)/void shader_main()/
