#version 460
#extension GL_EXT_ray_tracing : require
// Fixture variant of the 3-ray-debug closest-hit stage: the ids of the primary hit instead of its barycentrics,
// packed into the rgb8 render target: r = primitive[7:0], g = primitive[15:8], b = primitive[19:16] | instance << 4.
layout(location = 0) rayPayloadInEXT vec3 hitValue;
hitAttributeEXT vec3 attribs;

void main()
{
    const uint prim = uint(gl_PrimitiveID);
    const uint inst = uint(gl_InstanceID);
    const uvec3 packed = uvec3(prim & 255u, (prim >> 8) & 255u, ((prim >> 16) & 15u) | ((inst & 15u) << 4));
    hitValue = vec3(packed) / 255.0;
}
