"""Example transport for `kai0_b200.serving`: the per-connection loop (`serving.MessageHandler`) under the `websockets`
library, speaking the protocol of the reference's `WebsocketPolicyServer` (src/openpi/serving/websocket_policy_server.py:
metadata frame on connect, msgpack-numpy request / reply frames, a traceback text frame + close 1011 on failure, `/healthz`),
so the reference's own client (`openpi_client.websocket_client_policy.WebsocketClientPolicy`) talks to it unchanged.

Not part of the product (the socket is out of the hot path's scope): it exists so that tests/test_serving_cpu.py can put
the reference's client in front of this repo's serving stack, and as the glue a deployment would start from.

    python tools/serve_policy_b200.py --checkpoint ckpts/pi05_agilex/30000 --asset-id agilex \\
        --tokenizer paligemma_tokenizer.model --port 8000 [--max-batch 8 --max-wait-ms 2]
"""
from __future__ import annotations

import argparse
import asyncio
import http
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from kai0_b200 import serving  # noqa: E402


async def serve(policy, host: str = "0.0.0.0", port: int | None = 8000, metadata: dict | None = None, *, ready=None, stop=None):
    """Serve `policy` (anything with `infer(obs)`: a `serving.Policy`, or one `serving.RequestBatcher` shared by all
    connections) until `stop` (an asyncio.Event) is set.  `ready(port)` is called once the socket is bound."""
    import websockets.asyncio.server as ws_server
    import websockets.exceptions

    async def handler(ws):
        h = serving.MessageHandler(policy, metadata)
        await ws.send(h.greeting())
        loop = asyncio.get_running_loop()
        try:
            async for frame in ws:
                reply = await loop.run_in_executor(None, h.handle, frame)  # the model call must not block the event loop
                await ws.send(reply)
                if h.closed:
                    await ws.close(code=1011, reason="Internal server error. Traceback included in previous frame.")
                    break
        except websockets.exceptions.ConnectionClosed:
            pass

    def health(connection, request):
        if request.path == "/healthz":
            return connection.respond(http.HTTPStatus.OK, "OK\n")
        return None

    async with ws_server.serve(handler, host, port, compression=None, max_size=None, process_request=health) as server:
        if ready is not None:
            ready(server.sockets[0].getsockname()[1])
        if stop is None:
            await server.serve_forever()
        else:
            await stop.wait()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", required=True, help="a step directory: model.safetensors + assets/<asset-id>/norm_stats.json")
    ap.add_argument("--asset-id", required=True)
    ap.add_argument("--tokenizer", required=True, help="paligemma_tokenizer.model (SentencePiece)")
    ap.add_argument("--default-prompt", default=None)
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--max-batch", type=int, default=8)
    ap.add_argument("--max-wait-ms", type=float, default=2.0)
    args = ap.parse_args()
    from kai0_b200.pi0_pytorch import Pi05EngineConfig, PI0Pytorch

    cfg = Pi05EngineConfig()
    tok = serving.PaligemmaTokenizer(cfg.max_token_len, model_path=args.tokenizer)
    model = PI0Pytorch(cfg, max_batch=args.max_batch, init_weights=False)
    policy = serving.create_trained_policy(model, args.checkpoint, asset_id=args.asset_id, tokenizer=tok,
                                           default_prompt=args.default_prompt)
    with serving.RequestBatcher(policy, max_batch=args.max_batch, max_wait_ms=args.max_wait_ms) as batcher:
        asyncio.run(serve(batcher, args.host, args.port, policy.metadata))


if __name__ == "__main__":
    main()
