"""Synthetic VDL Mode 2 burst / IQ-stream generator (the inverse of the decoder).

The reference ships no modulator and no RS encoder; BASELINE.json's synthetic configurations
need one.  Everything here is derived from what the reference *decoder* accepts
(citations are reference file:line):

  AVLC frame + FCS (src/crc.c:21-64, src/avlc.c:39-40,177-179)
  -> HDLC flags + zero-bit insertion (inverse of src/bitstream.c:109-150)
  -> transmission length, RS(255,249) blocks (src/rs.c:27-49), octet interleave (inverse of src/decode.c:135-163)
  -> 25-bit header with 5 parity bits (src/decode.c:55-61,111-122)
  -> scrambler x^15+x+1, IV 0x6959 (src/decode.c:50, src/bitstream.c:94-107)
  -> Gray-coded differential 8-PSK at 10 500 sym/s after the 16-symbol preamble (src/demod.c:107-124,223,257-264)
  -> raised-cosine pulse (alpha 0.6), frequency shift to the channel offset, AWGN, cu8/cs16 quantisation
     (src/demod.c:349-365).

Pure numpy; used by tests/ and bench.py to make inputs.  It never touches the GPU path.
"""
import numpy as np

SYMBOL_RATE = 10500
SPS = 10
RS_K, RS_N = 249, 255
PREAMBLE_STEPS = np.array([0, 3, -3, 1, 1, 2, 0, 4, -3, 4, -2, 3, 1, -2, -3, 0])  # x pi/4, src/demod.c:107-124
GRAY = np.array([0, 1, 3, 2, 6, 7, 5, 4])            # src/demod.c:223  (phase step k -> bits)
GRAY_INV = np.argsort(GRAY)                         # bits -> phase step
_H_ROWS = [0x001FFF0, 0x07E1FE8, 0x18E61E4, 0x1B6A662, 0x0D3CAA1]   # src/decode.c:55-61

# ---- GF(256)/0x187 tables and the RS generator polynomial (src/libfec/init_rs.h:48-103) ----
_EXP = np.zeros(512, np.int64)
_LOG = np.zeros(256, np.int64)
_v = 1
for _e in range(255):
    _EXP[_e] = _v
    _LOG[_v] = _e
    _v <<= 1
    if _v & 0x100:
        _v ^= 0x187
_EXP[255:510] = _EXP[0:255]


def _gmul(a, b):
    return 0 if a == 0 or b == 0 else int(_EXP[_LOG[a] + _LOG[b]])


_G = [1]
for _i in range(6):                                   # prod (x + alpha^(120+i)), _G[k] = coeff of x^k
    _r = int(_EXP[(120 + _i) % 255])
    _G = [0] + _G
    for _k in range(len(_G) - 1):
        _G[_k] ^= _gmul(_G[_k + 1], _r)


def rs_parity(msg249):
    """6 parity octets of the systematic RS(255,249) codeword whose message is msg249 (highest order first)."""
    rem = [0] * 6
    for byte in msg249:
        fb = int(byte) ^ rem[0]
        rem = rem[1:] + [0]
        if fb:
            lf = int(_LOG[fb])
            for k in range(6):
                gk = _G[5 - k]
                if gk:
                    rem[k] ^= int(_EXP[lf + _LOG[gk]])
    return bytes(rem)


def crc16_x25(data, init=0xFFFF):
    crc = init
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ 0x8408 if crc & 1 else crc >> 1
    return crc


def avlc_frame(payload):
    """payload (addresses + control + info) -> frame with the 2-octet FCS appended, low byte first."""
    fcs = crc16_x25(payload) ^ 0xFFFF
    return bytes(payload) + bytes([fcs & 0xFF, fcs >> 8])


def random_avlc_frame(rng, n_octets):
    """A random frame of n_octets total (>= 11, FCS included)."""
    return avlc_frame(rng.integers(0, 256, n_octets - 2, dtype=np.uint8).tobytes())


def fec_octets_for(last_len):                         # src/decode.c:124-133
    return 0 if last_len < 3 else 2 if last_len < 31 else 4 if last_len < 68 else 6


def header_word(datalen_bits):
    rev = int(format(datalen_bits & 0x1FFFF, "017b")[::-1], 2)
    word = rev << 5
    for row, h in enumerate(_H_ROWS):
        word |= (bin(word & h & ~0x1F).count("1") & 1) << (4 - row)
    return word


_LFSR_SEQ = None


def scrambler_sequence(n):
    global _LFSR_SEQ
    if _LFSR_SEQ is None:
        s, out = 0x6959, np.zeros(32767, np.uint8)
        for i in range(32767):
            bit = (s ^ (s >> 14)) & 1
            s = (s >> 1) | (bit << 14)
            out[i] = bit
        _LFSR_SEQ = out
    reps = -(-n // 32767)
    return np.tile(_LFSR_SEQ, reps)[:n]


def hdlc_payload_bits(frames):
    """flag, stuffed frame, flag, stuffed frame, ..., flag  -> string of '0'/'1' (octets LSB first)."""
    flag = "01111110"
    s = flag
    for fr in frames:
        bits = "".join(format(b, "08b")[::-1] for b in fr)
        s += bits.replace("11111", "111110") + flag
    return s


def burst_bits(frames, corrupt_octets=None, header_bit_errors=()):
    """All bits that follow the preamble (scrambled), padded with zeros to a whole number of symbols.

    corrupt_octets: optional list of (block_row, column, xor_value) applied after RS encoding;
    header_bit_errors: bit positions (0 = first transmitted) flipped in the 25-bit header.
    Returns (bits uint8 array, info dict)."""
    return burst_bits_from_payload(hdlc_payload_bits(frames), corrupt_octets, header_bit_errors)


def burst_bits_from_payload(payload, corrupt_octets=None, header_bit_errors=()):
    """Same, from an arbitrary payload bit string ('0'/'1', transmission order): lets tests craft HDLC edge cases
    (repeated flags, aborts, frames that are not whole octets, missing closing flag)."""
    datalen = len(payload)
    if datalen > 0x3FFF:
        raise ValueError("burst too long")
    payload += "0" * (-datalen % 8)
    octets = np.array([int(payload[i:i + 8][::-1], 2) for i in range(0, len(payload), 8)], np.uint8)
    n_oct = len(octets)
    nblk = -(-n_oct // RS_K)
    last_len = n_oct - (nblk - 1) * RS_K
    table = np.zeros((nblk, RS_N), np.uint8)
    nfec = []
    for r in range(nblk):
        row = octets[r * RS_K:(r + 1) * RS_K]
        table[r, :len(row)] = row
        table[r, RS_K:] = np.frombuffer(rs_parity(table[r, :RS_K]), np.uint8)
        nfec.append(6 if r < nblk - 1 else fec_octets_for(last_len))
    if sum(nfec) == 0:
        raise ValueError("burst too short to carry FEC")
    for (r, c, x) in (corrupt_octets or []):
        table[r, c] ^= x
    # interleave: column-major over the data part, then over the FEC part (inverse of src/decode.c:135-163)
    tx = [table[r, c] for c in range(RS_K) for r in range(nblk) if not (r == nblk - 1 and c >= last_len)]
    tx += [table[r, RS_K + c] for c in range(6) for r in range(nblk) if c < nfec[r]]
    tx = np.array(tx, np.uint8)
    hdr = header_word(datalen)
    hbits = np.array([(hdr >> (24 - i)) & 1 for i in range(25)], np.uint8)
    for p in header_bit_errors:
        hbits[p] ^= 1
    dbits = np.unpackbits(tx, bitorder="little")
    bits = np.concatenate([hbits, dbits])
    bits = np.concatenate([bits, np.zeros(-len(bits) % 3, np.uint8)])
    bits ^= scrambler_sequence(len(bits))
    return bits, dict(datalen_bits=datalen, datalen_octets=n_oct, num_blocks=nblk, fec_octets=int(sum(nfec)),
                      header=hdr, n_symbols=len(bits) // 3)


def burst_phase_steps(bits):
    """Cumulative phase (in units of pi/4, integers) of every symbol: 16 preamble symbols then data."""
    tri = bits.reshape(-1, 3)
    sym_bits = tri[:, 0] * 4 + tri[:, 1] * 2 + tri[:, 2]          # MSB first, src/demod.c:274
    steps = GRAY_INV[sym_bits]
    data_phase = PREAMBLE_STEPS[-1] + np.cumsum(steps)
    return np.concatenate([PREAMBLE_STEPS, data_phase])


def rc_pulse(up, alpha=0.6, span=4):
    t = np.arange(-span * up, span * up + 1) / up
    num = np.sinc(t) * np.cos(np.pi * alpha * t)
    den = 1.0 - (2.0 * alpha * t) ** 2
    sing = np.abs(den) < 1e-9
    out = np.where(sing, np.pi / 4 * np.sinc(1.0 / (2.0 * alpha)), num / np.where(sing, 1.0, den))
    return out


def modulate_burst(bits, up, ramp_up=4, ramp_down=2):
    """Complex baseband at `up` samples per symbol, unit mean power over the burst body.
    Returns (samples complex64, index of the sample at the centre of the first preamble symbol)."""
    ph = burst_phase_steps(bits).astype(np.float64) * (np.pi / 4)
    amp = np.ones(len(ph))
    if ramp_up:
        ph = np.concatenate([np.full(ramp_up, ph[0]), ph])
        amp = np.concatenate([np.linspace(0.25, 1.0, ramp_up), amp])
    if ramp_down:
        ph = np.concatenate([ph, np.full(ramp_down, ph[-1])])
        amp = np.concatenate([amp, np.linspace(0.5, 0.0, ramp_down)])
    imp = np.zeros(len(ph) * up, np.complex128)
    imp[::up] = amp * np.exp(1j * ph)
    pulse = rc_pulse(up)
    n = len(imp) + len(pulse) - 1
    nfft = 1 << (n - 1).bit_length()
    sig = np.fft.ifft(np.fft.fft(imp, nfft) * np.fft.fft(pulse, nfft))[:n]
    first = (len(pulse) - 1) // 2 + ramp_up * up
    return sig.astype(np.complex64), first


class BurstSpec:
    """One burst to place in a stream."""
    def __init__(self, start_s, offset_hz, frames, power_dbfs=-20.0, freq_err_hz=0.0,
                 corrupt_octets=None, header_bit_errors=(), payload=None):
        self.start_s, self.offset_hz, self.frames = start_s, offset_hz, frames
        self.payload = payload          # raw payload bit string instead of frames (HDLC edge cases)
        self.power_dbfs, self.freq_err_hz = power_dbfs, freq_err_hz
        self.corrupt_octets, self.header_bit_errors = corrupt_octets, header_bit_errors
        self.info = None


def synth_stream(fs, duration_s, bursts, es_n0_db=None, noise_power=None, fmt="u8", seed=0):
    """Sum the bursts into one complex stream at `fs`, add AWGN and quantise.

    es_n0_db: if given, the full-band noise power is Ps*(fs/10500)/(Es/N0) with Ps the power of the
    *first* burst; noise_power overrides.  Returns a numpy array of uint8
    (interleaved I,Q cu8) or int16 (interleaved cs16)."""
    rng = np.random.default_rng(seed)
    up = fs // SYMBOL_RATE
    if up * SYMBOL_RATE != fs:
        raise ValueError("fs must be a multiple of 10500")
    n = int(round(duration_s * fs))
    x = np.zeros(n, np.complex64)
    for b in bursts:
        if b.payload is not None:
            bits, info = burst_bits_from_payload(b.payload, b.corrupt_octets, b.header_bit_errors)
        else:
            bits, info = burst_bits(b.frames, b.corrupt_octets, b.header_bit_errors)
        b.info = info
        sig, first = modulate_burst(bits, up)
        a = 10.0 ** (b.power_dbfs / 20.0)
        s0 = int(round(b.start_s * fs))
        k = np.arange(len(sig))
        rot = np.exp(2j * np.pi * (b.offset_hz + b.freq_err_hz) / fs * (k + s0)).astype(np.complex64)
        seg = (a * sig * rot).astype(np.complex64)
        lo, hi = max(s0, 0), min(s0 + len(seg), n)
        if hi > lo:
            x[lo:hi] += seg[lo - s0:hi - s0]
        b.info["first_symbol_sample"] = s0 + first
    if noise_power is None and es_n0_db is not None and bursts:
        ps = 10.0 ** (bursts[0].power_dbfs / 10.0)
        noise_power = ps * (fs / SYMBOL_RATE) / (10.0 ** (es_n0_db / 10.0))
    if noise_power:
        sd = np.sqrt(noise_power / 2.0)
        x += (rng.standard_normal(n, np.float32) * sd + 1j * rng.standard_normal(n, np.float32) * sd).astype(np.complex64)
    iq = np.empty(2 * n, np.float32)
    iq[0::2], iq[1::2] = x.real, x.imag
    if fmt == "u8":
        return np.clip(np.rint(iq * 127.5 + 127.5), 0, 255).astype(np.uint8)
    if fmt == "s16":
        return np.clip(np.rint(iq * 32768.0), -32768, 32767).astype(np.int16)
    raise ValueError(fmt)


def burst_duration_s(frames, ramp_up=4, ramp_down=2, payload=None):
    """On-air duration of the burst carrying `frames` (ramp + preamble + header/data symbols)."""
    _, info = burst_bits_from_payload(payload) if payload is not None else burst_bits(frames)
    return (ramp_up + len(PREAMBLE_STEPS) + info["n_symbols"] + ramp_down) / SYMBOL_RATE


def random_frames(rng, n_frames=None, lo=32, hi=240):
    n_frames = n_frames or int(rng.integers(1, 3))
    return [random_avlc_frame(rng, int(rng.integers(lo, hi + 1))) for _ in range(n_frames)]


def slot_offsets(n_slots=64, spacing_hz=25e3):
    """n_slots channel offsets on the 25 kHz raster, symmetric about (and excluding) the centre frequency."""
    half = n_slots // 2
    return [spacing_hz * k for k in range(-half, half + 1) if k != 0][:n_slots]


def traffic_stream(fs=2100000, duration_s=4.0, n_slots=64, bursts_per_s=2.0, es_n0_db=20.0, power_dbfs=-20.0,
                   seed=0x56444C33, fmt="u8"):
    """BASELINE.json config 3/5 traffic model: n_slots 25 kHz slots, Poisson burst arrivals
    per slot, 1-2 AVLC frames of 32-240 octets per burst.  Returns (iq array, slot offsets, bursts)."""
    rng = np.random.default_rng(seed)
    offs = slot_offsets(n_slots)
    bursts = []
    for o in offs:
        t = float(rng.exponential(1.0 / bursts_per_s))
        while True:
            frames = random_frames(rng)
            d = burst_duration_s(frames)
            if t + d > duration_s - 0.005:
                break
            bursts.append(BurstSpec(t, o, frames, power_dbfs=power_dbfs))
            t += d + 0.005 + float(rng.exponential(1.0 / bursts_per_s))
    iq = synth_stream(fs, duration_s, bursts, es_n0_db=es_n0_db, fmt=fmt, seed=seed + 1)
    return iq, offs, bursts
