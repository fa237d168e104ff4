// IK.h -- planar two-bone inverse kinematics (replaces src/IK.h:5-52).

// law of cosines gives the angle at the root; the joint is the goal direction scaled to L1 and
// rotated by that angle about z
SBX_FN vec3 ik_2_bone_centered_solver(_in(vec3) goal, _in(float) L1, _in(float) L2) {
    const float G = length(goal);
    const float cos_theta = (L1 * L1 + G * G - L2 * L2) / (2.0f * L1 * G);
    const float sin_theta = sqrt(1.0f - cos_theta * cos_theta);
    const mat3 rot = mat3(cos_theta, -sin_theta, 0.0f,
                          sin_theta, cos_theta, 0.0f,
                          0.0f, 0.0f, 1.0f);
    return rot * (normalize(goal) * L1);
}

SBX_FN vec3 ik_solver(_in(vec3) start, _in(vec3) goal, _in(float) bone_length_1, _in(float) bone_length_2) {
    return start + ik_2_bone_centered_solver(goal - start, bone_length_1, bone_length_2);
}
