// gate_resolve.inl -- host-side gate scalars, computed exactly as each reference *_apply does (see engine.h: GateK).
// Included inside namespace spz by abi.cu (the product) and by tests/emu/direct_emu.cpp (CPU emulation of the kernels).
int resolve_gate(int kind, const double *p, GateK *out) {
    out->kind = kind;
    for (double &v : out->s) v = 0.0;
    switch (kind) {
    case SPZ_GATE_H: case SPZ_GATE_X: case SPZ_GATE_Y: case SPZ_GATE_Z: return SPZ_OK;
    case SPZ_GATE_P: { // p_apply gates.rs:865: sin_cos(angle)
        out->s[0] = std::cos(p[0]); out->s[1] = std::sin(p[0]); return SPZ_OK; }
    case SPZ_GATE_RX: { // rx_apply gates.rs:749-751: theta = angle*0.5; ct = cos; nst = -sin
        const double th = p[0] * 0.5; out->s[0] = std::cos(th); out->s[1] = -std::sin(th); return SPZ_OK; }
    case SPZ_GATE_RY: { // ry_apply gates.rs:1125-1126: (sin, cos) of angle*0.5
        const double th = p[0] * 0.5; out->s[0] = std::sin(th); out->s[1] = std::cos(th); return SPZ_OK; }
    case SPZ_GATE_RZ: { // rz_apply gates.rs:970-973: d0 = (c, -s), d1 = (c, s)
        const double th = p[0] * 0.5; out->s[0] = std::cos(th); out->s[1] = std::sin(th); return SPZ_OK; }
    case SPZ_GATE_U: { // u_apply gates.rs:1286-1304
        const double st = std::sin(p[0] * 0.5), ct = std::cos(p[0] * 0.5);
        const double sl = std::sin(p[2]), cl = std::cos(p[2]);
        const double spl = std::sin(p[1] + p[2]), cpl = std::cos(p[1] + p[2]);
        const double sp = std::sin(p[1]), cp = std::cos(p[1]);
        out->s[0] = ct; out->s[1] = -cl * st; out->s[2] = -sl * st; out->s[3] = cp * st; out->s[4] = sp * st;
        out->s[5] = cpl * ct; out->s[6] = spl * ct;
        return SPZ_OK; }
    default:
        set_error("gate kind %d cannot be applied as a 2x2 pair update", kind);
        return SPZ_ERR_UNSUPPORTED;
    }
}
